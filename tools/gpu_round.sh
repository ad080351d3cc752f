set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-330
python tools/measure_configs.py > gpurun_out/configs_r1j.jsonl 2> gpurun_out/configs_r1j.err; tail -3 gpurun_out/configs_r1j.err
python - <<'PY'
import json
for line in open("gpurun_out/configs_r1j.jsonl"):
    r=json.loads(line)
    print(r["config"], {k:(round(v["ms"],3) if isinstance(v,dict) and "ms" in v else v) for k,v in r.items() if isinstance(v,dict)})
PY
