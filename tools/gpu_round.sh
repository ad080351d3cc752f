set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; cut -c1-400 gpurun_out/bench_r1m.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1m.json 2> gpurun_out/bench_ref_r1m.err; cut -c1-300 gpurun_out/bench_ref_r1m.json
python tools/measure_configs.py --cpu > gpurun_out/configs_r1m.jsonl 2> gpurun_out/configs_r1m.err; tail -3 gpurun_out/configs_r1m.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
