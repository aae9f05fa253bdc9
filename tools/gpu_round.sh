#!/bin/bash
# One GPU call: parity tests, bench line, ncu launch list, ncu full capture of the hot kernels.  tools/gpu_round.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_step.py --views 2 --train-batches 2 > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"march_kernel|app_forward_mma2_kernel|app_backward_mma_kernel|app_scatter_kernel|wgrad_mma_kernel|ray_backward_kernel" \
    -c 14 -o gpurun_out/${TAG}_full -f python tools/prof_step.py --views 1 --train-batches 1 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
