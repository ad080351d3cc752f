set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-330
python tools/measure_configs.py > gpurun_out/configs_r1e.jsonl 2> gpurun_out/configs_r1e.err; tail -3 gpurun_out/configs_r1e.err
python - <<'PY'
import json
for line in open("gpurun_out/configs_r1e.jsonl"):
    r=json.loads(line)
    print(r["config"], {k:(round(v["ms"],3) if isinstance(v,dict) and "ms" in v else v) for k,v in r.items() if isinstance(v,dict)})
PY
python tools/measure_configs.py --only cfg1 --iters 2 > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:filter_kernel_tc -s 2 -c 1 -o gpurun_out/prof_filter_tc_r1 python tools/measure_configs.py --only cfg1 --iters 2 > gpurun_out/ncu_full3.log 2>&1
