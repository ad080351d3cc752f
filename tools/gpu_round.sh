# one GPU round (1 GPU): the GPU test suite, the default bench line + reference arm, the ncu launch list of the bench
# command and ncu --set full captures of the dominant kernels (each ncu pass only after the plain command exited 0),
# the per-config and front-end timings.   bash tools/gpu_round.sh <tag>   -> gpurun_out/<tag>_*
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_multi.py > $OUT/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_gpu_tests.log
tail -4 $OUT/${TAG}_gpu_tests.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -c 400 $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
python tools/measure_configs.py --iters 5 --cpu > $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_bench.err
CBFRRT_GRID_MIN=12 python tools/measure_configs.py --iters 5 --only cfg1,cfg3 > $OUT/${TAG}_configs_gridmin12.jsonl 2>> $OUT/${TAG}_bench.err
python tools/measure_frontend.py > $OUT/${TAG}_frontend.jsonl 2>> $OUT/${TAG}_bench.err
python tools/measure_planner_sized.py > $OUT/${TAG}_planner_sized.jsonl 2>> $OUT/${TAG}_bench.err
python tools/eval_planning.py > $OUT/${TAG}_planning.jsonl 2>> $OUT/${TAG}_bench.err
python tools/measure_lidar.py --states 256 > $OUT/${TAG}_lidar.jsonl 2>> $OUT/${TAG}_bench.err
python tools/measure_lidar.py --states 2048 --rays 64 --cloud 512 >> $OUT/${TAG}_lidar.jsonl 2>> $OUT/${TAG}_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --scale 0.125"
$B > $OUT/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collide_kernel -s 6 -c 1 -o $OUT/${TAG}_ncu_full_geometry_pass $B > $OUT/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:filter_kernel_tc -s 6 -c 1 -o $OUT/${TAG}_ncu_full_network_pass $B >> $OUT/${TAG}_ncu_full.log 2>&1
CBFRRT_SPLIT=0 ncu --set full --clock-control none --import-source on -k regex:safety_kernel_tc -s 6 -c 1 -o $OUT/${TAG}_ncu_full_fused $B >> $OUT/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:backbone_kernel_v2 -s 2 -c 1 -o $OUT/${TAG}_ncu_full_lidar python tools/measure_lidar.py --states 64 --iters 2 > $OUT/${TAG}_ncu_full_lidar.log 2>&1
ls -la $OUT/${TAG}_ncu_full*.ncu-rep
