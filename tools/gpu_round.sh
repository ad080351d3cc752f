# one GPU round: the GPU test suite, the bench line, per-config timings, smoke
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-260 gpurun_out/bench_final.json
python tools/measure_configs.py --cpu > gpurun_out/configs_final.jsonl 2> gpurun_out/configs_final.err; tail -2 gpurun_out/configs_final.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
