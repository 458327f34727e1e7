#!/bin/bash
# Round-1 closing GPU visit: parity suite, smoke, both bench arms, the (f)-row tools, ncu launch list and two
# --set full captures.  Outputs under gpurun_out/ (copied into profiles/ by hand).
TAG=${1:-r1f}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | grep -v -i warn | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench fp32 B=32"; timeout 600 python bench.py --warmup 3 --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
echo "== bench bf16 B=32"; timeout 600 python bench.py --warmup 3 --precision bf16 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_$TAG.json | python tools/bench_summary.py
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep -v -i warn | tee $OUT/bench_ref_$TAG.json | python tools/bench_summary.py
echo "== per-frame latency"; timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -8 | tee $OUT/latency_$TAG.log
echo "== keypoint heads"; timeout 200 python tools/bench_kp.py 2>&1 | grep -v -i warn | tail -6 | tee $OUT/kp_$TAG.log
echo "== AT_net2 / clip"; timeout 200 python tools/bench_at.py 300 1 2>&1 | grep -v -i warn | tee $OUT/at_$TAG.log | head -3
timeout 200 python tools/clip_e2e.py 300 --no-oracle 2>&1 | grep -v -i warn | tee $OUT/clip_$TAG.log
KREGEX='regex:conv_tc_kernel|conv_simt_kernel|aa_downsample_kernel|kp_stage_kernel|flow_combine_kernel|warp_occlude|warp_image_kernel|nchw_to_act_kernel|pack_image_kernel'
echo "== ncu launch list (2 steps after 3 warm-up steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 105 -c 70 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1; tail -1 $OUT/ncu_bench_$TAG.log | cut -c1-160
echo "== ncu --set full: one bottleneck conv, warp_occlude, one split-K hourglass conv"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 103 -c 1 -f -o $OUT/prof_conv_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; tail -1 $OUT/ncu_full_$TAG.log | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_occlude -s 3 -c 1 -f -o $OUT/prof_warp_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full2_$TAG.log 2>&1; tail -1 $OUT/ncu_full2_$TAG.log | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 94 -c 1 -f -o $OUT/prof_enc4_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full3_$TAG.log 2>&1; tail -1 $OUT/ncu_full3_$TAG.log | cut -c1-160
