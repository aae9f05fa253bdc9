#!/bin/bash
# Quick GPU iteration: appearance-related parity tests, cycle trace, forward/backward timing (no CPU baseline).
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/trace_mma.py 2>&1 | tail -5
timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("fwd", d["value"], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
if d.get("fwd_bwd"): print("fwd_bwd", d["fwd_bwd"]["value"], d["fwd_bwd"]["kernel_ms"])
P
