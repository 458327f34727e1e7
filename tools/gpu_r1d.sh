#!/bin/bash
# Round-1 fourth GPU visit: new defaults (kx-in-N 112, split-K, 32-byte stores) -- parity suite, bench lines,
# role counters, ncu captures of the `first` and `final` convs.
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== pytest -m gpu"
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_r1d.log
echo "== bench fp32 B=32"
timeout 300 python bench.py --warmup 3 --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_r1d.json | python tools/bench_summary.py
echo "== bench bf16 B=32"
timeout 300 python bench.py --warmup 3 --no-cpu-baseline --all-kernels --precision bf16 2>&1 | grep -v -i warn | tee $OUT/bench_bf16_r1d.json | python tools/bench_summary.py
echo "== role counters"
EAMM_TC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep tc_prof | tail -29 | tee $OUT/tc_prof_r1d.log | cut -c1-260 | head -16
echo "== ncu first / final"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 87 -c 1 -o $OUT/prof_first_r1d -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_first.log 2>&1; tail -2 $OUT/ncu_first.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 115 -c 1 -o $OUT/prof_final_r1d -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_final.log 2>&1; tail -2 $OUT/ncu_final.log | cut -c1-200
