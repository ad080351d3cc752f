set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; cut -c1-260 gpurun_out/bench_r1z.json
python tools/measure_configs.py --cpu > gpurun_out/configs_r1z.jsonl 2> gpurun_out/configs_r1z.err; tail -2 gpurun_out/configs_r1z.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
