set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/measure_configs.py --only cfg2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l); print(r['config'], {k:round(v['ms'],3) for k,v in r.items() if isinstance(v,dict) and 'ms' in v})"
