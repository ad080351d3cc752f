# one GPU round: ncu --set full capture of the bench kernel (after a plain run), then the GPU test suite
set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:safety_kernel_tc -s 40 -c 1 -o gpurun_out/prof_safety_tc_r1e python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_e.log 2>&1
ls -la gpurun_out/prof_safety_tc_r1e.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
