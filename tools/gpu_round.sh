set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; cut -c1-400 gpurun_out/bench_r1i.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1i.json 2> gpurun_out/bench_ref_r1i.err; cut -c1-400 gpurun_out/bench_ref_r1i.json
python tools/measure_configs.py --cpu > gpurun_out/configs_r1i.jsonl 2> gpurun_out/configs_r1i.err; tail -3 gpurun_out/configs_r1i.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
