#!/bin/bash
# One GPU-box visit: parity tests, stage error tables, benchmark lines, ncu launch list.
# Usage: bash tools/gpu_round.sh [tag]   (outputs under gpurun_out/)
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv unit checks"; timeout 900 python tools/gpu_conv_check.py ${ONLY:+--only $ONLY} 2>&1 | grep -v -i warn | tail -25 | tee $OUT/conv_$TAG.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
echo "== stage check full fp32"; timeout 300 python tools/gpu_stage_check.py --config full --batch 2 --precision fp32 2>&1 | grep -v -i warn | tee $OUT/stage_full_fp32_$TAG.log
echo "== stage check full bf16"; timeout 300 python tools/gpu_stage_check.py --config full --batch 2 --precision bf16 2>&1 | grep -v -i warn | tee $OUT/stage_full_bf16_$TAG.log
echo "== bench fp32 B=32"; timeout 600 python bench.py --warmup 3 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json
echo "== bench bf16 B=32"; timeout 600 python bench.py --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_$TAG.json
echo "== AT_net2 per-clip time, config-5 clip (no oracle leg)"; timeout 200 python tools/bench_at.py 300 1 2>&1 | grep -v -i warn | tee $OUT/at_$TAG.log | head -3
timeout 200 python tools/clip_e2e.py 300 --no-oracle 2>&1 | grep -v -i warn | tee $OUT/clip_$TAG.log
if [ -n "$QUICK" ]; then exit 0; fi
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | grep -v -i warn | tee $OUT/bench_ref_$TAG.json
KREGEX='regex:conv_tc_kernel|conv_simt_kernel|aa_downsample_kernel|kp_stage_kernel|flow_combine_kernel|warp_occlude|warp_image_kernel|nchw_to_act_kernel|pack_image_kernel'
echo "== ncu launch list (our kernels, 2 steps after 3 warm-up steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 108 -c 72 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1; tail -2 $OUT/ncu_bench_$TAG.log | cut -c1-200
echo "== ncu --set full: one bottleneck conv + warp_occlude"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 103 -c 1 -o $OUT/prof_conv_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; tail -2 $OUT/ncu_full_$TAG.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_occlude -s 3 -c 1 -o $OUT/prof_warp_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full2_$TAG.log 2>&1; tail -2 $OUT/ncu_full2_$TAG.log | cut -c1-200
