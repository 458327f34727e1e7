#!/bin/bash
# GPU visit: conv cases (filter), role counters of one forward, bench fp32.
TAG=${1:-r2n}; ONLY=${2:-"ah ,mix,f16,cta2,res"}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv check"; timeout 900 python tools/gpu_conv_check.py --only "$ONLY" > $OUT/conv_$TAG.log 2>&1; grep -c "^OK" $OUT/conv_$TAG.log; grep -v "^OK" $OUT/conv_$TAG.log | tail -25
echo "== roles"; EAMM_TC_PROF=2 timeout 300 python tools/prof_step.py 2>&1 | grep tc_prof | cut -c1-330 > $OUT/roles_$TAG.log; tail -29 $OUT/roles_$TAG.log | cut -c9-330
AB="--steps 20 --warmup 3 --no-cpu-baseline --no-extras --all-kernels"
echo "== bench fp32"; timeout 600 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
