set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
python tools/measure_configs.py --cpu > gpurun_out/configs_r1.jsonl 2> gpurun_out/configs_r1.err; tail -3 gpurun_out/configs_r1.err; cat gpurun_out/configs_r1.jsonl
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:safety_kernel -s 3 -c 2 -o gpurun_out/prof_safety_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
