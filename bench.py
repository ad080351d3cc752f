#!/usr/bin/env python3
"""bench.py -- headline benchmark of the CBF-INC-RRT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg1|cfg0|cfg2|cfg3|cfg4]

A "step" = one pass of the fused hot path (FK -> link/obstacle distances + masks -> min-distance observation and
gradient -> barrier MLP + state-Jacobian -> CBF-QP) over one batch of synthetic states.  Default workload =
BASELINE.json configs[1]: Magician 4-DoF, 16 box obstacles + floor, 2^20 sampled states per GPU (weak scaling;
rank r takes the seeded sample r).  Prints ONE JSON line (rank 0).

  value        states/s with inputs resident in HBM (CUDA events around the launches of the timed steps, L2 flushed
               between steps, max over ranks); for N > 1 the step includes the NCCL all-gather of the validity
               masks and filtered controls (block-cyclic, overlapped on a second stream).
  e2e          the same metric through the host-buffer C-ABI call (cbfrrt_safety_filter_host): pinned host
               buffers, H2D of q and D2H of (safe, u) inside the timed region.
  roofline     the path is FP32 CUDA-core bound: achieved algorithmic TFLOP/s of the fused kernel
               against the FMA rate measured in this run by cbfrrt_fma_probe; the HBM fraction (algorithmic bytes
               against MEASURED_PEAKS.json) is reported beside it.
  cpu_baseline the oracle port of the reference's CPU path on this box's host cores, bounded sample.
  --impl reference   times only that CPU port (the reference itself cannot run: pybullet / SCS are absent).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "collision-checked + CBF-QP-filtered states/sec (fused safety filter)"
UNIT = "states/s"


# ----------------------------------------------------------------------------- algorithmic cost model
def algorithmic_cost(robot, n_boxes, n_spheres, floor, outputs="all"):
    """SURVEY.md 8d formulas -> (flops per state, HBM bytes per state) for the fused q -> (masks, datax, h, J, u, r)."""
    from cbfrrt._robot_tables import ROBOT_MODELS
    m = ROBOT_MODELS[robot]
    n, F, S = m["n"], m["n_links"], len(m["spheres"])
    fk = 78 * F
    centres = 15 * S
    pairs = S * (22 * n_boxes + 10 * n_spheres + 3 * int(floor))
    self_pairs = 10 * len(m["self_sphere_pairs"])
    grad = 12 * n
    mlp = 2 * 2 * ((n + 1) * 48 + 2 * 48 * 48 + 48)   # forward + reverse mode
    qp = 2 * n * (3 * n + 4)
    flops = fk + centres + pairs + self_pairs + grad + mlp + qp
    if outputs == "all":
        byts = 4 * n + 1 + 1 + 4 * (2 * n + 1) + 4 + 4 * n + 4 * n + 4
    else:  # valid + u
        byts = 4 * n + 1 + 4 * n
    return flops, byts


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                mx = float(f[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU port (oracle) leg
def cpu_port_run(w, rows, threads):
    from oracle import c_oracle
    c_oracle.set_threads(threads)
    q = w["q"][:rows]
    t0 = time.perf_counter()
    c_oracle.safety_filter(w["robot"], q, w["boxes"], w["spheres"], True, w["weights"], w["hidden"], w["layers"],
                           w["goal"], w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
    return rows / (time.perf_counter() - t0)


def cpu_sequential_run(w, rows=2048):
    """The reference's loop SHAPE (SURVEY 8d "reference-faithful sequential mode"): one call per state, single thread --
    arm_dynamics.py:199-201 / :274-276 loop over the batch in Python and clf_controller solves one QP per row.  Here each
    iteration is one oracle call on a batch of 1 (exact QP), so this prices the per-state call overhead of that structure
    around a C core, not pybullet / SCS themselves (absent here)."""
    from oracle import c_oracle
    c_oracle.set_threads(1)
    rows = int(min(rows, w["q"].shape[0]))
    q = w["q"]
    args_ = (w["boxes"], w["spheres"], True, w["weights"], w["hidden"], w["layers"], w["goal"], w["K"], w["u_lo"], w["u_hi"],
             w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
    for i in range(min(64, rows)):
        c_oracle.safety_filter(w["robot"], q[i:i + 1], *args_)
    t0 = time.perf_counter()
    for i in range(rows):
        c_oracle.safety_filter(w["robot"], q[i:i + 1], *args_)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{rows} states of {w['name']}, one call per state (per-state Python loop around the C port)"}


def reference_arm(args, w):
    """--impl reference: the CPU port of the reference's path (oracle/cbfrrt_oracle.c, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = min(w["q"].shape[0], 1 << 18)
    cpu_port_run(w, min(rows, 8192), threads)  # warm caches / page in the library
    vals = []
    for _ in range(max(1, args.warmup)):
        cpu_port_run(w, rows, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_port_run(w, rows, threads))
    dt = time.perf_counter() - t0
    v = rows * args.steps / dt
    sample = f"{rows} of {w['q'].shape[0]} states of {w['name']} per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "note": "reference cannot run here (pybullet/cvxpylayers/SCS absent); "
                   "this is the repo's C port of its CPU path (oracle/cbfrrt_oracle.c) on all host threads"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "sequential": cpu_sequential_run(w)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: pin this rank's host threads (and so its first-touch pinned buffers of the e2e leg) to the CPUs
    NVML reports as local to its GPU, so that the host<->device copies of 8 ranks do not all cross the socket link."""
    if os.environ.get("CBFRRT_BENCH_NO_AFFINITY", "0") == "1":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception as e:   # no NVML / restricted container: keep the default affinity
        sys.stderr.write(f"[bench] cpu affinity not set ({e!r})\n")
    return None


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cbfrrt")
    ap.add_argument("--workload", default="cfg1")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    from cbfrrt import workloads as W
    cfg_idx = {"cfg0": 0, "cfg1": 1, "cfg4": 4}[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, W.config(cfg_idx, scale=args.scale))
        return

    import torch
    import torch.distributed as dist
    from cbfrrt import ops
    from cbfrrt.sharded import ShardedSweep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = W.config(cfg_idx, scale=args.scale)
    robot, rid, n = w["robot"], ops.ROBOT_ID[w["robot"]], w["n"]
    if w["q"] is None:   # cfg4: sampled on the device, shard-invariant
        B = w["n_states"] // world if cfg_idx == 4 else w["n_states"]
        q = ops.sample_states(B, w["sample_seed"], rank * B, rid, device=dev)
    else:
        B = w["q"].shape[0]
        q_np = w["q"] if rank == 0 else W.sample_states(robot, B, 1000 + rank)   # weak scaling: own sample per rank
        q = torch.as_tensor(q_np, device=dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
    gr = ops.build_grid(b8, s4) if b8.shape[0] + s4.shape[0] > 0 else ops.no_grid(dev)
    net, goal, K, lo, hi = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev), t(w["goal"]), t(w["K"]), t(w["u_lo"]), t(w["u_hi"])
    common = (goal, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    if world == 1:
        def step():
            return ops.safety_filter(q, *common)
        launches_per_step, outputs = 1, "all"
    else:
        outputs = "valid+u"
        try:   # fused epilogue all-gather over NVLink peer mappings (one kernel + one barrier per step)
            from cbfrrt.sharded import FusedSweep
            sweep = FusedSweep(lambda qb, vp, up, off: ops.valid_control_scatter(qb, *common, vp, up, off), n, B, dev)
            def step():
                return sweep(q)
            launches_per_step = 1
            collective = ("fused: kernel epilogue stores (valid u8, u f32xn) into every rank's replica over NVLink ("
                          + ("one multicast store per row, replicated by NVSwitch" if sweep.multicast else "per-peer stores")
                          + ", symmetric memory) + barrier")
        except Exception as e:   # no P2P / symmetric memory: NCCL all-gather, block-cyclic, overlapped on a 2nd stream
            sys.stderr.write(f"[bench] symmetric memory unavailable ({e!r}); using NCCL all_gather\n")
            n_blocks = 4
            block = B // n_blocks
            sweep = ShardedSweep(lambda qb: ops.valid_control(qb, *common), n, block, device=dev)
            def step():
                return sweep(q[:block * n_blocks])
            launches_per_step, collective = n_blocks, "nccl all_gather(valid u8, u f32xn), block-cyclic, overlapped"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clock ramp: a fresh process starts with the GPU idle, and a handful of sub-millisecond warm-up steps is too short
    # for the SM clock to reach its sustained value -- run the step untimed for about 0.3 s first, then the W warm-up steps
    # (a fixed count, identical on every rank: the multi-GPU step ends in a barrier)
    for _ in range(min(20000, max(50, int(0.3 / (B * 4.5e-10))))):
        step()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = ops.launch_count()
    barrier()
    wall0 = time.time()
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        flush.zero_()           # L2 flush between timed iterations (not inside the event pair)
    barrier()
    wall1 = time.time()
    launches = ops.launch_count() - l0
    clocks = sampler.stop(wall0, wall1)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    units_per_step = B * world
    value = units_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant (only) kernel: launch duration measured above (one launch per step at N=1)
    flops_u, bytes_u = algorithmic_cost(robot, w["boxes"].shape[0], w["spheres"].shape[0], True, outputs)
    sink = torch.zeros(1, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    grid, blockdim, iters = sms * 8, 256, 1 << 15
    for _ in range(2):
        ops.lib.cbfrrt_fma_probe(grid, blockdim, iters, sink.data_ptr(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = torch.cuda.current_stream().cuda_stream
    e0.record()
    for _ in range(3):
        ops.lib.cbfrrt_fma_probe(grid, blockdim, iters, sink.data_ptr(), stream_ptr)
    e1.record()
    torch.cuda.synchronize()
    fp32_peak = 3 * grid * blockdim * iters * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    kernel_s = ms * 1e-3 / args.steps
    ach_tf = flops_u * B / kernel_s / 1e12
    ach_gbs = bytes_u * B / kernel_s / 1e9
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured"
    except Exception:
        peaks, hbm_peak, hbm_src = {}, 6650.0, "fallback"
    tensor_peak = next((float(v) for k, v in peaks.items() if "bf16" in k.lower() and isinstance(v, (int, float))), None)
    tensor_src = "measured (dense bf16, MEASURED_PEAKS.json)" if tensor_peak else "fallback (nominal dense bf16)"
    tensor_peak = tensor_peak or 2250.0
    # DRAM traffic of one launch from the committed ncu --set full capture of this exact command
    # (profiles/r1e_ncu_full_safety_kernel_tc_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum); only valid for the
    # default workload at N=1, else null
    traffic = 45.3e6 if (world == 1 and cfg_idx == 1 and B == 1 << 20) else None
    mlp_flops = 2 * 2 * ((n + 1) * 48 + 2 * 48 * 48 + 48)
    roofline = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                "traffic": traffic, "peak_source": "cbfrrt_fma_probe in this run (FP32 FMA; not in MEASURED_PEAKS.json)",
                "flops_per_state": flops_u, "bytes_per_state": bytes_u,
                "frac_without_mlp": (flops_u - mlp_flops) * B / kernel_s / 1e12 / fp32_peak,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "peak_source": hbm_src},
                "tensor": {"achieved": 3 * mlp_flops * B / kernel_s / 1e12, "peak": tensor_peak, "unit": "TFLOP/s",
                           "frac": 3 * mlp_flops * B / kernel_s / 1e12 / tensor_peak, "peak_source": tensor_src,
                           "note": "executed tcgen05 flops of the barrier network (3-term TF32 split = 3x algorithmic)"},
                "kernel": "safety_kernel_tc" if world == 1 else "safety_kernel_tc (valid+u, fused exchange)",
                "note": "achieved = ALGORITHMIC flops of SURVEY 8d (brute-force pair sweep + FP32 MLP) / kernel time, "
                        "against the CUDA-core FP32 peak; the kernel runs the MLP on tcgen05 (TF32x3) and culls pairs "
                        "exactly, so frac may exceed the CUDA-core bound; frac_without_mlp counts only the geometry + QP",
                "kernel_ms": kernel_s * 1e3}

    # ---- end to end through the host-buffer C-ABI (pinned host buffers; H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        q_host = torch.empty((B, n), dtype=torch.float32).pin_memory()
        q_host.copy_(q.cpu())
        safe_h = torch.empty(B, dtype=torch.uint8).pin_memory()
        u_h = torch.empty((B, n), dtype=torch.float32).pin_memory()
        out = {"safe": safe_h.numpy(), "u": u_h.numpy()}
        def e2e_step():
            ops.safety_filter_host(robot, q_host.numpy(), w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"],
                                   w["K"], w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"],
                                   out, device=local_rank)
        for _ in range(args.warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": units_per_step * args.steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(B * n * 4), "d2h_bytes_per_step": int(B * (1 + 4 * n)),
               "api": "cbfrrt_safety_filter_host (C-ABI, pinned host q -> safe,u)"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and w["q"] is not None:
        threads = os.cpu_count() or 1
        rows = min(B, 1 << 17)
        cpu_port_run(w, 4096, threads)
        vals = [cpu_port_run(w, rows, threads) for _ in range(3)]
        cpu = {"value": float(np.median(vals)), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{rows} of {B} states of {w['name']}, median of 3 (oracle/cbfrrt_oracle.c, pthreads)",
               "sequential": cpu_sequential_run(w)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "robot": robot, "states_per_gpu": B, "boxes": int(w["boxes"].shape[0]),
                       "spheres": int(w["spheres"].shape[0]), "floor": True, "mlp": "(n+1)-48-48-48-1",
                       "outputs": outputs, "l2": "flushed by a 256 MiB write between timed steps (outside the event pairs)",
                       "collective": "none" if world == 1 else collective,
                       "host_cpus_bound": affinity},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
