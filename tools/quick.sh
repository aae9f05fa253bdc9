#!/bin/bash
# Quick GPU iteration: parity tests, forward/backward kernel timing (no CPU baseline).  tools/quick.sh TAG [pytest -k expr]
set -u
mkdir -p gpurun_out
TAG=${1:-q}
KEXPR=${2:-}
if [ -n "$KEXPR" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -4; else timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("fwd", round(d["value"],3), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
if d.get("fwd_bwd"): print("fwd_bwd", round(d["fwd_bwd"]["value"],4), {k: round(v,3) for k,v in d["fwd_bwd"]["kernel_ms"].items()}, d["fwd_bwd"]["full_iteration"]["fused_ms_per_iteration"])
P
