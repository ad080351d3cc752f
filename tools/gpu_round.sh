set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/measure_configs.py --only cfg3 > gpurun_out/configs_r1q.jsonl 2> gpurun_out/configs_r1q.err; tail -3 gpurun_out/configs_r1q.err
python - <<'PY'
import json
for line in open("gpurun_out/configs_r1q.jsonl"):
    r=json.loads(line)
    print(r["config"], {k:(round(v["ms"],3) if isinstance(v,dict) and "ms" in v else v) for k,v in r.items() if isinstance(v,dict)})
PY
