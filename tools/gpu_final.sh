#!/bin/bash
# Round-closing GPU visit: the evidence set in one go (tests, smoke, bench lines of both arms, next-row tools, role
# counters, ncu launch list + summarised full captures).   usage: bash tools/gpu_final.sh TAG [full]
#   "full" adds every conv unit case and the fp16 B=256 ncu pass (about 5 more minutes)
TAG=${1:-r2z}; FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$FULL" ]; then QUICK=1 ONLY=all bash tools/gpu_round.sh $TAG; else QUICK=1 bash tools/gpu_round.sh $TAG; fi
exec </dev/null
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep -v -i warn | tee $OUT/bench_ref_$TAG.json | python tools/bench_summary.py
echo "== role counters of the production kernels (EAMM_TC_PROF=2)"
EAMM_TC_PROF=2 timeout 300 python tools/prof_step.py 2>&1 | grep tc_prof | tail -29 | cut -c1-340 > $OUT/roles_$TAG.log; wc -l $OUT/roles_$TAG.log
if [ -n "$FULL" ]; then
  bash tools/gpu_ncu.sh $TAG fp32 32 "first:0 down0:1 down1:2 res_conv1:14 res_conv2:15 up0:26 up1:27 final:28"
  bash tools/gpu_ncu.sh ${TAG} fp16 256 "down0:1 res_conv1:14 res_conv2:15"
else
  bash tools/gpu_ncu.sh $TAG fp32 32 "first:0 res_conv1:14 res_conv2:15 up1:27"
fi
