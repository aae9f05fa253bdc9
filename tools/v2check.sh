#!/bin/bash
# Quick GPU check of the forward appearance kernel: tensor-core parity tests, then forward timing with both kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-v2}
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or steady or bench_shape or 64cube or edge or view_heads" 2>&1 | tail -15
echo "pytest rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("fwd", round(d["value"],3), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
    if d.get("fwd_bwd"): print("fwd_bwd", round(d["fwd_bwd"]["value"],4), {k: round(v,3) for k,v in d["fwd_bwd"]["kernel_ms"].items()})
    print("parity", d.get("parity"))
except Exception as e:
    print("no bench line", e)
P
