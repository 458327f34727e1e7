#!/bin/bash
# Round-1 third GPU visit: split-K for the small hourglass maps and 256-bit epilogue stores (both opt-in here).
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv unit checks: split-K"
timeout 400 python tools/gpu_conv_check.py --only splitk 2>&1 | grep -v -i warn | tail -9 | tee $OUT/conv_splitk.log
echo "== conv unit checks: 256-bit stores"
EAMM_TC_ST256=1 timeout 400 python tools/gpu_conv_check.py --only "cta2,pairfold,first,up2 256,kxw 7x7 128->16 64x64 N=5" 2>&1 | grep -v -i warn | tail -16 | tee $OUT/conv_st256.log
for cfg in "" "EAMM_TC_SPLITK=1" "EAMM_TC_ST256=1" "EAMM_TC_SPLITK=1 EAMM_TC_ST256=1"; do
  echo "== bench fp32 B=32 [$cfg]"
  env $cfg timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee "$OUT/bench_fp32_${cfg// /_}.json" | python tools/bench_summary.py
done
echo "== bench bf16 B=32 [both]"
env EAMM_TC_SPLITK=1 EAMM_TC_ST256=1 timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels --precision bf16 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_both.json | python tools/bench_summary.py
echo "== pytest -m gpu with both"
env EAMM_TC_SPLITK=1 EAMM_TC_ST256=1 timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_both.log
