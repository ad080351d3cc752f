#!/usr/bin/env python3
"""What the host link of this box can carry with N ranks copying at once: per rank, pinned host <-> device copies of
256 MiB in both directions concurrently (two streams), with and without binding the rank to its GPU's NUMA node before
the pinned allocation.  Prints per-rank and aggregate GB/s (the ceiling of bench.py's e2e leg).
    python -m torch.distributed.run --nproc-per-node N tools/experiments/pcie_probe.py"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa_node  # noqa: E402


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = {}
    for bind in (False, True):
        if bind:
            res["cpus_bound"] = bind_to_gpu_numa_node(lr)
        nbytes = 256 << 20
        h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
        h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
        d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        d_out = torch.ones(nbytes, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for mode in ("h2d", "d2h", "both"):
            for it in range(2):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(8):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            d_in.copy_(h_in, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            gbs = 8 * nbytes * (2 if mode == "both" else 1) / dt / 1e9
            t = torch.tensor([gbs], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t)
            res[f"{mode}_{'bound' if bind else 'unbound'}"] = {"rank_gbs": gbs, "aggregate_gbs": float(t.item())}
        del h_in, h_out
    if rank == 0:
        print(json.dumps({"world": world, **res}))
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
            print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
        except Exception as e:
            print(e)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
