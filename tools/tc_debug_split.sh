for d in 0 1 2; do echo "== EAMM_TC_DEBUG=$d"; EAMM_TC_DEBUG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -v -i warn | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step', round(d['ms_per_step'],3)); print({k:v['ms'] for k,v in d['kernels_ms_per_step'].items()})
"; done
