set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; cut -c1-600 gpurun_out/bench_r1f.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
