#!/bin/bash
# GPU visit: conv unit cases (optionally filtered), bench A/B of the halo-tile scheme, GPU test suite.
TAG=${1:-r2h}; ONLY=${2:-}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv check $ONLY"; timeout 1200 python tools/gpu_conv_check.py ${ONLY:+--only "$ONLY"} > $OUT/conv_$TAG.log 2>&1; grep -c "^OK" $OUT/conv_$TAG.log; grep -v "^OK" $OUT/conv_$TAG.log | tail -25
AB="--steps 20 --warmup 3 --no-cpu-baseline --no-extras --all-kernels"
echo "== bench fp32 (halo tiles)"; timeout 600 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
echo "== bench fp32 EAMM_TC_AH=0"; EAMM_TC_AH=0 timeout 600 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_noah_$TAG.json | python tools/bench_summary.py
echo "== bench fp16"; timeout 600 python bench.py $AB --precision fp16 2>&1 | grep -v -i warn | tee $OUT/bench_fp16_$TAG.json | python tools/bench_summary.py
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -q -m gpu -x > $OUT/pytest_$TAG.log 2>&1; tail -15 $OUT/pytest_$TAG.log
