#!/bin/bash
# Round-1 second GPU visit: validate and time the opt-in conv_tc variants (wide folded pair step, 112-column
# kx-in-N schemes) against the defaults.  Outputs under gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
NEW="EAMM_TC_KXW=3 EAMM_TC_CTA2=11"
echo "== conv unit checks (new variants)"
timeout 400 python tools/gpu_conv_check.py --only kxw 2>&1 | grep -v -i warn | tail -12 | tee $OUT/conv_kxw.log
timeout 300 python tools/gpu_conv_check.py --only pfwide 2>&1 | grep -v -i warn | tail -8 | tee $OUT/conv_pfwide.log
echo "== bench fp32 B=32 default"
timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_default.json | python tools/bench_summary.py
echo "== bench fp32 B=32 new variants"
env $NEW timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_new.json | python tools/bench_summary.py
echo "== bench bf16 B=32 new variants"
env $NEW timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels --precision bf16 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_new.json | python tools/bench_summary.py
echo "== pytest -m gpu with the new variants"
env $NEW timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_new.log
echo "== per-role cycle counters (single-CTA instrumented kernel), last forward"
EAMM_TC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep tc_prof | tail -30 | tee $OUT/tc_prof_default.log | cut -c1-250
