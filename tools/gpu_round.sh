# one GPU round (1 GPU): the GPU test suite, the default bench line + reference arm, the ncu launch list of the bench
# command and one ncu --set full capture of its dominant kernel (each ncu pass only after the plain command exited 0).
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>   -> gpurun_out/<tag>_*
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_multi.py > $OUT/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_gpu_tests.log
tail -5 $OUT/${TAG}_gpu_tests.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --scale 0.125 > $OUT/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --scale 0.125 > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:safety_kernel_tc -s 6 -c 1 -o $OUT/${TAG}_ncu_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --scale 0.125 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_ncu_full.ncu-rep
