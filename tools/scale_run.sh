# strong-scaling sweep of bench.py (default workload: BASELINE config 5) on one box: N in $NS (default "1 2 4 8");
# one JSON per N under gpurun_out/<tag>_scale_<N>.json.   bash tools/scale_run.sh <tag>
TAG=${1:-r2}
for N in ${NS:-1 2 4 8}; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_scale_$N.json 2> gpurun_out/${TAG}_scale_$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_scale_$N.json 2> gpurun_out/${TAG}_scale_$N.err; fi
  tail -1 gpurun_out/${TAG}_scale_$N.json | python -c 'import json,sys; r=json.loads(sys.stdin.read()); print(r["n_gpus"], "value %.4g" % r["value"], "ms %.3f" % r["ms_per_step"], "e2e %.4g" % (r["e2e"] or {}).get("value", 0), r["config"]["collective"][:110])' || tail -5 gpurun_out/${TAG}_scale_$N.err
done
