"""Recipe for ``oracle/_ref``: the reference's OWN implementation of the hot path, made runnable on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/t2n_oracle.py's header for who may import oracle/).

The reference's path is three pure-PyTorch files with no dependency beyond torch + numpy
(models/tensorBase.py, models/tensoRF.py, models/sh.py).  ``/root/reference`` does not exist on the GPU
box, so -- exactly like a compiled reference's ``.so`` -- this recipe places a runnable artefact under
``oracle/_ref/`` (git-ignored: it never enters the history; NOT gpurun-ignored: it travels with the snapshot).  The
files are taken verbatim from where they lie under the reference checkout at build time; nothing is edited.

``bench.py --impl reference`` and the ``cpu_baseline`` leg time ``oracle/_ref`` when it is present
(``cpu_baseline.kind == "reference"``) and fall back to the oracle port (``"port"``) otherwise.

    python oracle/make_ref.py [--ref /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("models/tensorBase.py", "models/tensoRF.py", "models/sh.py")


def make(ref="/root/reference", out=os.path.join(HERE, "_ref")) -> bool:
    """Returns True when oracle/_ref is in place (built now or already there), False when the checkout is absent."""
    if not os.path.isdir(ref):
        return os.path.exists(os.path.join(out, "MANIFEST.json"))
    os.makedirs(os.path.join(out, "models"), exist_ok=True)
    manifest = {"source": ref, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(ref, rel), os.path.join(out, rel)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(out, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    ok = make(a.ref)
    print("oracle/_ref", "ready" if ok else "NOT built (no reference checkout)")
    sys.exit(0 if ok else 1)
