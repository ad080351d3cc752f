set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "filter or safety or full_size or rollout or golden" 2>&1 | tail -2
python tools/measure_configs.py --only cfg1,cfg3,cfg4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l); print(r['config'], {k:round(v['ms'],3) for k,v in r.items() if isinstance(v,dict) and 'ms' in v})"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-240
