#!/usr/bin/env python
"""Static SASS opcode histogram of the built library per kernel (cuobjdump -sass): the mnemonics that prove tcgen05 / TMEM /
TMA use (UTCHMMA, LDTM, STTM, UTCBAR, UBLKCP, UTMALDG) next to the atomics (RED, ATOM, ATOMS) and plain loads.

    python tools/sass_histogram.py [text2nerf_b200/lib/libt2n_b200.so] > profiles/r2_sass_histogram.md
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "text2nerf_b200/lib/libt2n_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS", "LDG", "STG",
        "LDS", "STS", "SHFL", "FFMA", "MUFU", "HMMA"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur]["total"] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("REDG", "ATOMG") and op == k):
                hist[cur][k] += 1
                break
        else:
            if op.startswith("UTC") and op not in KEYS:
                hist[cur]["UTC*other"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except OSError:
        return n


cols = ["total", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS", "LDG", "STG", "LDS", "STS", "SHFL", "FFMA", "MUFU"]
print("# Static SASS opcode histogram of `%s` (cuobjdump -sass, sm_100a)\n" % lib)
print("UTCHMMA = tcgen05.mma (kind::tf32); LDTM / STTM = tcgen05.ld / st (TMEM); UTCBAR = tcgen05.commit; UBLKCP = cp.async.bulk (1-D TMA);")
print("UTMALDG = tensor-map TMA (none: every bulk copy here is a contiguous pre-swizzled image); REDG / RED = red.global; ATOMS = shared atomics.\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for fn, h in hist.items():
    name = demangle(fn)
    if not name.startswith("void t2n::") and not name.startswith("t2n::"):
        continue
    print("| `" + name.replace("void ", "") + "` | " + " | ".join(str(h.get(c, 0)) for c in cols) + " |")
