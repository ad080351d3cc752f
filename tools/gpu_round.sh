set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; cut -c1-300 gpurun_out/bench_r1o.json
