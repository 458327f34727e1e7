#!/bin/bash
# split-K cost model: B=32 plans must be unchanged, batch-1 latency back to (or better than) the unsplit numbers
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv unit checks: split-K"
timeout 400 python tools/gpu_conv_check.py --only splitk 2>&1 | grep -v -i warn | tail -9 | tee $OUT/conv_splitk_r1g.log
echo "== per-frame latency"; timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -4 | tee $OUT/latency_r1g.log
echo "== per-frame latency, split-K off"; EAMM_TC_SPLITK=0 timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -4 | tee $OUT/latency_r1g_nosplit.log
echo "== bench fp32 B=32"; timeout 600 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_r1g.json | python tools/bench_summary.py
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_r1g.log
