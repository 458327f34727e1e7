#!/bin/bash
# One GPU-box visit: parity tests, stage error tables, benchmark lines, ncu launch list.
# Usage: bash tools/gpu_round.sh [tag]   (outputs under gpurun_out/)
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== stage check full fp32"; timeout 300 python tools/gpu_stage_check.py --config full --batch 2 --precision fp32 2>&1 | grep -v -i warn | tee $OUT/stage_full_fp32_$TAG.log
echo "== stage check full bf16"; timeout 300 python tools/gpu_stage_check.py --config full --batch 2 --precision bf16 2>&1 | grep -v -i warn | tee $OUT/stage_full_bf16_$TAG.log
echo "== bench fp32 B=32"; timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json
echo "== bench bf16 B=32"; timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_$TAG.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | grep -v -i warn | tee $OUT/bench_ref_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 80 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1; tail -3 $OUT/ncu_bench_$TAG.log | cut -c1-300
