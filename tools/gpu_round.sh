set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -2 gpurun_out/bench_r1d.err; cut -c1-330 gpurun_out/bench_r1d.json
python tools/measure_configs.py --cpu > gpurun_out/configs_r1d.jsonl 2> gpurun_out/configs_r1d.err; tail -3 gpurun_out/configs_r1d.err
python - <<'PY'
import json
for line in open("gpurun_out/configs_r1d.jsonl"):
    r=json.loads(line)
    print(r["config"], {k:(round(v["ms"],3) if isinstance(v,dict) and "ms" in v else v) for k,v in r.items() if isinstance(v,dict)})
PY
