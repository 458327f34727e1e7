#!/bin/bash
# split-K with 256-bit partial traffic, S <= 9 and the refined cost model: conv checks, batch-1 latency, B=32 bench
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv unit checks: split-K"
timeout 400 python tools/gpu_conv_check.py --only splitk 2>&1 | grep -v -i warn | tail -9 | tee $OUT/conv_splitk_r1h.log
echo "== per-frame latency (S <= 9)"; timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -2 | tee $OUT/latency_r1h.log
echo "== per-frame latency (S <= 4)"; EAMM_TC_SPLITK_MAX=4 timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -2 | tee $OUT/latency_r1h_max4.log
echo "== bench fp32 B=32"; timeout 600 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_r1h.json | python tools/bench_summary.py
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_r1h.log
