#!/bin/bash
# GPU visit: halo-tile conv cases, bench (fp32 / AH off for N = 256 / fp16), ncu summaries of the narrow layers.
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv check"; timeout 900 python tools/gpu_conv_check.py --only "ah ,mix,f16" > $OUT/conv_$TAG.log 2>&1; grep -c "^OK" $OUT/conv_$TAG.log; grep -v "^OK" $OUT/conv_$TAG.log | tail -25
AB="--steps 20 --warmup 3 --no-cpu-baseline --no-extras --all-kernels"
echo "== bench fp32"; timeout 600 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
echo "== bench fp32 EAMM_TC_AH=2 (no halo tiles for N = 256)"; EAMM_TC_AH=2 timeout 600 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_ah2_$TAG.json | python tools/bench_summary.py
echo "== bench fp16"; timeout 600 python bench.py $AB --precision fp16 2>&1 | grep -v -i warn | tee $OUT/bench_fp16_$TAG.json | python tools/bench_summary.py
bash tools/gpu_ncu.sh $TAG fp32 32 "first:0 down0:1 down1:2 res_conv1:14 res_conv2:15 up0:26 up1:27"
