set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/experiments/small_batch_sweep.py 2>&1 | grep -E " (1024|4096|8192|16384) "
CBFRRT_COMPACT=0 python tools/experiments/small_batch_sweep.py 2>&1 | grep -E " (1024|4096|8192) " | sed 's/^/nocompact /'
python tools/measure_configs.py --only cfg0 2>/dev/null | python -c "import json,sys; r=json.loads(sys.stdin.readline()); print(r['config'], {k:round(v['ms'],4) for k,v in r.items() if isinstance(v,dict) and 'ms' in v})"
