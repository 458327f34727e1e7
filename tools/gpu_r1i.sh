#!/bin/bash
# last visit of the session: cost-model N tile for under-filled layers (EAMM_TC_BNCOST) against the older rule
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== pytest -m gpu (default)"; timeout 300 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_r1i.log
echo "== bench fp32 B=32 (default, full line)"; timeout 300 python bench.py --warmup 3 --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_r1i.json | python tools/bench_summary.py
echo "== latency BNCOST=1"; timeout 100 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -2 | tee $OUT/latency_r1i.log
echo "== latency BNCOST=0"; EAMM_TC_BNCOST=0 timeout 100 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -2 | tee $OUT/latency_r1i_bn0.log
echo "== bench fp32 B=32 BNCOST=0"; EAMM_TC_BNCOST=0 timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_r1i_bn0.json | python tools/bench_summary.py
