#!/bin/bash
# GPU visit: MMA step-cost micro-benchmark, error-by-scope table of the mixed scheme, full GPU test log, smoke.
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== mma step cost"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/mma_step_cost.cu -o /tmp/mma_step_cost && timeout 120 /tmp/mma_step_cost 2>&1 | tee $OUT/mma_step_cost_$TAG.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -q -m gpu -s > $OUT/pytest_full_$TAG.log 2>&1; grep -v -i warn $OUT/pytest_full_$TAG.log | tail -60
grep -h "max-abs\|PSNR\|batch invariance\|configs\[" $OUT/pytest_full_$TAG.log > $OUT/pytest_errors_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | grep -v -i warn | tail -6 | tee $OUT/smoke_$TAG.log
echo "== mixed-scheme error by scope"; timeout 900 python tools/gpu_mix_error.py 2>&1 | grep -v -i warn | tee $OUT/mix_error_$TAG.txt
