set -x
python tools/experiments/small_one.py > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:safety_kernel -s 4 -c 1 -o gpurun_out/prof_small_r1 python tools/experiments/small_one.py > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out/prof_small_r1.ncu-rep
