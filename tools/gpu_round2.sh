#!/bin/bash
# Round-2 GPU visit: conv unit checks of the fp16 / mixed cases, the GPU test suite, the benchmark line with its
# extra legs, A/B switches of the mixed scheme and the per-role cycle counters.   usage: bash tools/gpu_round2.sh TAG
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
exec </dev/null
echo "== conv unit checks (f16, mix)"
timeout 600 python tools/gpu_conv_check.py --only "f16,mix" 2>&1 | grep -v -i warn | tail -40 | tee $OUT/conv_$TAG.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -q -m gpu -s 2>&1 | grep -v -i warn | tail -120 | tee $OUT/pytest_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | grep -v -i warn | tail -6 | tee $OUT/smoke_$TAG.log
echo "== bench fp32 B=32 (all legs)"; timeout 900 python bench.py --warmup 3 --all-kernels 2>&1 | grep -v -i warn | tee $OUT/bench_fp32_$TAG.json | python tools/bench_summary.py
if [ -n "$QUICK" ]; then exit 0; fi
AB="--warmup 3 --steps 20 --no-cpu-baseline --no-extras --all-kernels"
echo "== A/B: no pairs for narrow mixed layers (EAMM_TC_CTA2=3)"; EAMM_TC_CTA2=3 timeout 300 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_ab_cta2_3_$TAG.json | python tools/bench_summary.py
echo "== A/B: no cin=64 mixed variant (EAMM_B200_MIX64=0)"; EAMM_B200_MIX64=0 timeout 300 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_ab_mix64_0_$TAG.json | python tools/bench_summary.py
echo "== A/B: mixed scheme in the bottleneck only (EAMM_B200_MIX=res)"; EAMM_B200_MIX=res timeout 300 python bench.py $AB 2>&1 | grep -v -i warn | tee $OUT/bench_ab_mixres_$TAG.json | python tools/bench_summary.py
echo "== A/B: 3-pass bf16 everywhere (fp32_bf16x3)"; timeout 300 python bench.py $AB --precision fp32_bf16x3 2>&1 | grep -v -i warn | tee $OUT/bench_ab_bf16x3_$TAG.json | python tools/bench_summary.py
echo "== per-role cycle counters (single-CTA instrumented kernel), last forward"
EAMM_TC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep tc_prof | tail -29 > $OUT/tc_prof_$TAG.log; wc -l $OUT/tc_prof_$TAG.log
echo "== per-frame latency"; timeout 200 python tools/bench_latency.py 2>&1 | grep -v -i warn | tail -4 | tee $OUT/latency_$TAG.log
