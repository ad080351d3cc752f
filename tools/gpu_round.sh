set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:safety_kernel_tc -s 4 -c 1 -o gpurun_out/prof_safety_tc_r1k python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_k.log 2>&1
tail -1 gpurun_out/plain_b.log | cut -c1-300
