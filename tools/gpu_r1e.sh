#!/bin/bash
# Round-1 fifth GPU visit: compact scheme-3 epilogue buffer (4 stages), 16-channel warp_occlude kernel (opt-in).
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv unit checks: kx-in-N 112"
timeout 400 python tools/gpu_conv_check.py --only kxw 2>&1 | grep -v -i warn | tail -9 | tee $OUT/conv_kxw_r1e.log
echo "== pytest -m gpu (EAMM_WARP16=1)"
EAMM_WARP16=1 timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_r1e.log
for cfg in "EAMM_WARP16=0" "EAMM_WARP16=1"; do
  echo "== bench fp32 B=32 [$cfg]"
  env $cfg timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee "$OUT/bench_fp32_r1e_${cfg}.json" | python tools/bench_summary.py
done
echo "== bench bf16 B=32 [EAMM_WARP16=1]"
EAMM_WARP16=1 timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels --precision bf16 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_r1e.json | python tools/bench_summary.py
