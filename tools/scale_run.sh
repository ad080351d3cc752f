# weak-scaling sweep of bench.py on one box: N = 1 2 4 8 (or the list in $NS); one JSON per N under gpurun_out/
for N in ${NS:-1 2 4 8}; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; fi
  tail -1 gpurun_out/scale_$N.json | python -c 'import json,sys; r=json.loads(sys.stdin.read()); print(r["n_gpus"], r["value"], r["ms_per_step"], r["e2e"] and r["e2e"]["value"], r["config"]["collective"][:100])'
done
