#!/bin/bash
# One GPU-box visit: parity tests, smoke, benchmark lines (both arms), the (f)-row tools, batch-1 latency,
# ncu launch list and --set full captures.  Usage: bash tools/gpu_round.sh [tag]   (outputs under gpurun_out/)
#   ONLY=substr[,substr]  also run tools/gpu_conv_check.py on the matching cases ("" = skip, "all" = every case, ~3 min)
#   QUICK=1               stop before the reference arm and the ncu passes
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
if [ -n "$ONLY" ]; then
  echo "== conv unit checks"
  if [ "$ONLY" = all ]; then timeout 900 python tools/gpu_conv_check.py 2>&1 | grep -v -i warn | tail -60 | tee $OUT/conv_$TAG.log
  else timeout 900 python tools/gpu_conv_check.py --only "$ONLY" 2>&1 | grep -v -i warn | tail -30 | tee $OUT/conv_$TAG.log; fi
fi
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests/ -q -m gpu -s 2>&1 | grep -v -i warn | tail -80 | tee $OUT/pytest_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | grep -v -i warn | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench fp32 B=32"; timeout 600 python bench.py --warmup 3 --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
echo "== bench fp16 B=32"; timeout 600 python bench.py --warmup 3 --precision fp16 --no-cpu-baseline --no-extras --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp16_$TAG.json | python tools/bench_summary.py
echo "== per-frame latency"; timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -4 | tee $OUT/latency_$TAG.log
echo "== keypoint heads"; timeout 200 python tools/bench_kp.py 2>&1 | grep -v -i warn | tail -6 | tee $OUT/kp_$TAG.log
echo "== AT_net2 per-clip time, config-5 clip (no oracle leg)"; timeout 200 python tools/bench_at.py 300 1 2>&1 | grep -v -i warn | tee $OUT/at_$TAG.log | head -3
timeout 200 python tools/clip_e2e.py 300 --no-oracle 2>&1 | grep -v -i warn | tee $OUT/clip_$TAG.log
if [ -n "$QUICK" ]; then exit 0; fi
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep -v -i warn | tee $OUT/bench_ref_$TAG.json | python tools/bench_summary.py
echo "== per-role cycle counters (single-CTA instrumented kernel), last forward"
EAMM_TC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep tc_prof | tail -29 > $OUT/tc_prof_$TAG.log; wc -l $OUT/tc_prof_$TAG.log
# launch order of one forward: pack_image, first, down0, down1, aa, kp_stage, hourglass x10, mask_occ, flow_combine,
# warp_occlude, warp_image, res x12, up0, up1, final = 35 launches, 29 of them conv_tc_kernel (first = 0, enc4 = 7,
# mask_occ = 13, res0.conv1 = 14, final = 28); 3 warm-up forwards precede the timed ones
KREGEX='regex:conv_tc_kernel|conv_simt_kernel|aa_downsample_kernel|kp_stage_kernel|flow_combine_kernel|warp_occlude|warp_image_kernel|nchw_to_act_kernel|pack_image_kernel'
echo "== ncu launch list (our kernels, 2 steps after 3 warm-up steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 105 -c 70 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1; tail -1 $OUT/ncu_bench_$TAG.log | cut -c1-160
echo "== ncu --set full: bottleneck conv, warp_occlude, first, final, split-K hourglass conv"
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k "$1" -s $2 -c 1 -f -o $OUT/prof_$3_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_$3_$TAG.log 2>&1; tail -1 $OUT/ncu_$3_$TAG.log | cut -c1-160; }
cap regex:conv_tc_kernel 103 conv
cap regex:warp_occlude 3 warp
cap regex:conv_tc_kernel 87 first
cap regex:conv_tc_kernel 115 final
cap regex:conv_tc_kernel 94 enc4
echo "summaries: python tools/ncu_summary.py $OUT/prof_*_$TAG.ncu-rep"
