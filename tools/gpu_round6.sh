#!/bin/bash
# GPU visit: all conv unit cases, bench fp32 + fp16, GPU test suite, ncu summaries of selected launches.
TAG=${1:-r2k}; CAPS=${2:-"down0:1 up1:27"}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv check"; timeout 1200 python tools/gpu_conv_check.py > $OUT/conv_$TAG.log 2>&1; grep -c "^OK" $OUT/conv_$TAG.log; grep -v "^OK" $OUT/conv_$TAG.log | tail -25
AB="--steps 20 --warmup 3 --no-cpu-baseline --no-extras --all-kernels"
echo "== bench fp32"; timeout 600 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
echo "== bench fp16"; timeout 600 python bench.py $AB --precision fp16 2>&1 | grep -v -i warn | tee $OUT/bench_fp16_$TAG.json | python tools/bench_summary.py
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -q -m gpu -x > $OUT/pytest_$TAG.log 2>&1; tail -5 $OUT/pytest_$TAG.log
bash tools/gpu_ncu.sh $TAG fp32 32 "$CAPS"
