#!/usr/bin/env python3
"""bench.py -- headline benchmark of the CBF-INC-RRT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4|cfg1|cfg0]

A "step" = one pass of the fused hot path (FK -> link/obstacle distances + masks -> min-distance observation and
gradient -> barrier MLP + state-Jacobian -> CBF-QP) over one batch of synthetic states.

Default workload = BASELINE.json configs[4] (the config the metric and the north-star targets are quoted on): Franka
Panda, 256 box obstacles + floor, 2^26 = 64 M sampled states, outputs `valid` (u8) and the filtered control `u`
(f32 x 7) per state.  STRONG scaling: the 64 M states are split over the N ranks (contiguous blocks, sampled on the
device by the counter-based sampler, so the result does not depend on N); for N > 1 every rank ends the step with the
full [64 M] mask and [64 M, 7] controls (exchange fused into the kernel epilogue over NVLink, or NCCL all-gather).
`--workload cfg1` (Magician, 16 boxes, 2^20 states per GPU, weak scaling, all outputs) and `cfg0` are kept.
Prints ONE JSON line (rank 0).

  value        states/s with inputs resident in HBM (CUDA events around the launches of the timed steps, max over ranks)
  e2e          the same metric through the host-buffer C-ABI call (cbfrrt_safety_filter_host): pinned host buffers,
               H2D of q and D2H of (valid, u) inside the timed region
  roofline     per pipe, never mixed: FP32 CUDA-core work against the FMA rate measured in this run
               (cbfrrt_fma_probe), tcgen05 work against a dense TF32 matmul measured in this run, HBM bytes against
               MEASURED_PEAKS.json; `frac` is the largest of the three and `bound` names it
  cpu_baseline the oracle port of the reference's CPU path on this box's host cores, bounded sample
  --impl reference   times only that CPU port (the reference itself cannot run: pybullet / SCS are absent); it does not
               load the CUDA library
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "cbf-inc-rrt_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "collision-checked + CBF-QP-filtered states/sec (fused safety filter)"
UNIT = "states/s"
CFG = {"cfg0": 0, "cfg1": 1, "cfg4": 4}
TARGETS_8GPU = {"collision_checked_franka_configs_per_s": 1e9, "cbf_qp_filtered_control_steps_per_s": 1e8}


def load_workloads(native: bool):
    """cbfrrt.workloads.  native=False (reference arm): through a stub parent package, so that neither
    cbfrrt/__init__.py nor libcbfrrt.so is loaded -- the module itself is numpy / torch-CPU only."""
    if native:
        from cbfrrt import workloads
        return workloads
    stub = types.ModuleType("cbfrrt")
    stub.__path__ = [os.path.join(PKG, "cbfrrt")]
    sys.modules["cbfrrt"] = stub
    return importlib.import_module("cbfrrt.workloads")


# ----------------------------------------------------------------------------- algorithmic cost model
def algorithmic_cost(W, robot, n_boxes, n_spheres, floor, outputs, pairs_per_state=None, lookups_per_state=0):
    """-> dict of algorithmic work per state, one entry per pipe.

    fp32: FK 78/frame + sphere centres 15/sphere + pair evaluations 22/box 10/sphere 3/floor + self pairs 10 + gradient
    12 n + QP 2n(3n+4)  (SURVEY.md 8d).  `pairs_per_state` = exact pair evaluations the broad phase leaves (measured on
    the device for this scene); None = brute force, every link sphere against every obstacle.  `lookups_per_state` =
    broad-phase look-ups (index arithmetic ~12 flop each).  The barrier network (forward + reverse mode) is counted on
    the tensor pipe when the tcgen05 kernels run it: executed flops = 3 x algorithmic (TF32 3-term split).
    hbm: compulsory bytes: q in, outputs out; scene / weights / grid are read once per CTA."""
    m = W.ROBOT_MODELS[robot]
    n, F, S = m["n"], m["n_links"], len(m["spheres"])
    brute = S * (22 * n_boxes + 10 * n_spheres)
    pair_flops = brute if pairs_per_state is None else 22.0 * pairs_per_state
    geom = 78 * F + 15 * S + pair_flops + 3 * S * int(floor) + 12 * lookups_per_state \
        + 10 * len(m["self_sphere_pairs"]) + 12 * n
    qp = 2 * n * (3 * n + 4)
    mlp = 2 * 2 * ((n + 1) * 48 + 2 * 48 * 48 + 48)   # forward + reverse mode
    if outputs == "all":
        byts = 4 * n + 1 + 1 + 4 * (2 * n + 1) + 4 + 4 * n + 4 * n + 4
    else:  # valid + u
        byts = 4 * n + 1 + 4 * n
    return {"fp32": geom + qp, "fp32_brute_force": 78 * F + 15 * S + brute + 3 * S * int(floor)
            + 10 * len(m["self_sphere_pairs"]) + 12 * n + qp, "mlp": mlp, "bytes": byts}


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
def host_states(w, rows):
    """the first `rows` states of the workload on the host (config 5: the oracle's copy of the counter-based sampler,
    bit-identical to the device sampler)"""
    if w["q"] is not None:
        return w["q"][:rows]
    from oracle import c_oracle
    return c_oracle.sample_states(w["robot"], w["sample_seed"], 0, rows)


def cpu_port_run(w, q, threads):
    from oracle import c_oracle
    c_oracle.set_threads(threads)
    t0 = time.perf_counter()
    c_oracle.safety_filter(w["robot"], q, w["boxes"], w["spheres"], True, w["weights"], w["hidden"], w["layers"],
                           w["goal"], w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
    return q.shape[0] / (time.perf_counter() - t0)


def cpu_sequential_run(w, rows=1024):
    """The reference's loop SHAPE (SURVEY 8d "reference-faithful sequential mode"): one call per state, single thread --
    arm_dynamics.py:199-201 / :274-276 loop over the batch in Python and clf_controller solves one QP per row.  Here each
    iteration is one oracle call on a batch of 1 (exact QP), so this prices the per-state call overhead of that structure
    around a C core, not pybullet / SCS themselves (absent here)."""
    from oracle import c_oracle
    c_oracle.set_threads(1)
    q = host_states(w, rows)
    rows = q.shape[0]
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


def cpu_sample_rows(w):
    """bounded CPU sample: about 1-2 s of work per pass on a 16-thread host (the 256-box brute-force sweep of the oracle
    costs ~200 kflop per state)"""
    total = w["q"].shape[0] if w["q"] is not None else w["n_states"]
    return int(min(total, 1 << 18 if w["boxes"].shape[0] <= 32 else 1 << 16))


def workload_config(w, extra=None):
    total = w["q"].shape[0] if w["q"] is not None else w["n_states"]
    c = {"workload": w["name"], "robot": w["robot"], "states": int(total), "boxes": int(w["boxes"].shape[0]),
         "spheres": int(w["spheres"].shape[0]), "floor": True, "obstacles": w["obstacles"], "mlp": "(n+1)-48-48-48-1"}
    c.update(extra or {})
    return c


def reference_arm(args):
    """--impl reference: the CPU port of the reference's path (oracle/cbfrrt_oracle.c, all host threads).  The CUDA
    library is not loaded by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = load_workloads(native=False)
    w = W.config(CFG[args.workload], scale=args.scale)
    threads = os.cpu_count() or 1
    q = host_states(w, cpu_sample_rows(w))
    rows = q.shape[0]
    cpu_port_run(w, q[:min(rows, 4096)], threads)  # page in the library
    for _ in range(max(1, args.warmup)):
        cpu_port_run(w, q, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_run(w, q, threads)
    dt = time.perf_counter() - t0
    v = rows * args.steps / dt
    total = w["q"].shape[0] if w["q"] is not None else w["n_states"]
    sample = f"the first {rows} of {total} states of {w['name']} per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if CFG[args.workload] == 4 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, {"outputs": "valid+u" if CFG[args.workload] == 4 else "all",
                                       "note": "reference cannot run here (pybullet/cvxpylayers/SCS absent); this is the repo's "
                                               "C port of its CPU path (oracle/cbfrrt_oracle.c) on all host threads; per-state "
                                               "rate on a bounded sample"}),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "sequential": cpu_sequential_run(w, 256 if w["boxes"].shape[0] > 32 else 1024)},
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


def timed_events(torch, fn, iters, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
        flush.zero_()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def other_configs(torch, ops, W, dev, flush, iters=3):
    """device-resident timings of the other BASELINE configs (N = 1 only; extra keys, not the headline)"""
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    out = {}
    for idx in (1, 2, 3):
        w = W.config(idx)
        rid, n = ops.ROBOT_ID[w["robot"]], w["n"]
        b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
        gr = ops.build_grid(b8, s4)
        net, K, lo, hi, goal = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev), t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), t(w["goal"])
        res = {"obstacles": w["obstacles"]}
        if idx == 1:
            q = t(w["q"])
            fn = lambda: ops.safety_filter(q, goal, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
            fn(); fn()
            ms = float(np.median(timed_events(torch, fn, iters, flush)))
            res.update(ms=ms, states_per_s=q.shape[0] / (ms * 1e-3), outputs="all")
        elif idx == 2:
            qf, qt = t(w["q"]), t(w["q_to"])
            fn = lambda: ops.edge_validity(qf, qt, w["n_interp"], b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)
            fn()
            ms = float(np.median(timed_events(torch, fn, iters, flush)))
            v = fn()[0]
            res.update(ms=ms, edges_per_s=qf.shape[0] / (ms * 1e-3), configs_per_s=qf.shape[0] * (w["n_interp"] + 1) / (ms * 1e-3),
                       valid_fraction=float(v.float().mean().item()))
        else:
            q0 = t(w["q"])
            T = q0.shape[0]
            safe = ops.masks(q0, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)[0].bool()   # "filtered to safe" 
            keep = q0[safe]
            q0 = keep.repeat((T + keep.shape[0] - 1) // keep.shape[0], 1)[:T].contiguous()
            fn = lambda: ops.rollout(q0, goal, w["steps"], w["ctrl_every"], w["dt"], 0, 10, 0.1, False, True, b8, s4, 1, gr, net, K, lo, hi,
                                     w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
            fn()
            ms = float(np.median(timed_events(torch, fn, iters, flush)))
            res.update(ms=ms, sim_steps_per_s=T * w["steps"] / (ms * 1e-3), qp_steps_per_s=T * w["steps"] / w["ctrl_every"] / (ms * 1e-3),
                       safe_start_fraction=float(safe.float().mean().item()))
        out[w["name"]] = res
        del w
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cbfrrt")
    ap.add_argument("--workload", default="cfg4", choices=sorted(CFG))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    cfg_idx = CFG[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    W = load_workloads(native=True)
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
    strong = cfg_idx == 4
    if strong:   # config 5: G states in total, split over the ranks; sampled on the device, shard-invariant
        G = (w["n_states"] // (world * 128)) * world * 128
        B = G // world
        q = ops.sample_states(B, w["sample_seed"], rank * B, rid, device=dev)
        outputs = "valid+u"
    else:        # weak scaling: every rank its own seeded sample of the config's size
        B = w["q"].shape[0]
        G = B * world
        q_np = w["q"] if rank == 0 else W.sample_states(robot, B, 1000 + rank)
        q = torch.as_tensor(q_np, device=dev)
        outputs = "all" if world == 1 else "valid+u"
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
    gr = ops.build_grid(b8, s4) if b8.shape[0] + s4.shape[0] > 0 else ops.no_grid(dev)
    net, goal, K, lo, hi = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev), t(w["goal"]), t(w["K"]), t(w["u_lo"]), t(w["u_hi"])
    common = (goal, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    collective = "none"
    if world == 1:
        if outputs == "all":
            def step():
                return ops.safety_filter(q, *common)
        else:
            valid_out = torch.empty(B, dtype=torch.uint8, device=dev)
            u_out = torch.empty((B, n), dtype=torch.float32, device=dev)
            def step():
                return ops.valid_control_into(q, *common, valid_out, u_out)
        launches_per_step = 1
    else:
        try:   # fused epilogue all-gather over NVLink peer mappings (one kernel + barriers per step)
            from cbfrrt.sharded import FusedSweep
            env_chunks = os.environ.get("CBFRRT_EXCHANGE_CHUNKS")
            sweep = FusedSweep(lambda qb, vp, up, off, mc: ops.valid_control_scatter(qb, *common, vp, up, off, mc), n, B, dev,
                               passes=(lambda qb, s_out, d_out: ops.observe_into(qb, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid, s_out, d_out),
                                       lambda dx, sv, rows, vp, up, off, mc: ops.cbf_filter_scatter(dx, sv, rows, goal, net, K, lo, hi, w["lam"],
                                                                                                   w["penalty"], 48, 2, n, vp, up, off, mc)),
                               chunks=int(env_chunks) if env_chunks else None)
            def step():
                return sweep(q)
            launches_per_step = 1
            collective = ("fused: kernel epilogue stores (valid u8, u f32xn) into every rank's replica over NVLink ("
                          + ("multimem.st to the NVSwitch multicast address, replicated by the switch" if sweep.multicast else "per-peer stores")
                          + ", symmetric memory) + barrier"
                          + (f"; pipelined in {sweep.chunks} chunks over two streams (stores of chunk k drain under the geometry pass of chunk k+1)"
                             if sweep.chunks > 1 else ""))
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

    # clock ramp: a fresh process starts with the GPU idle, and a handful of short warm-up steps is too short for the SM
    # clock to reach its sustained value -- run the step untimed for about 0.5 s first, then the W warm-up steps
    # (a fixed count, identical on every rank: the multi-GPU step ends in a barrier)
    for _ in range(int(min(20000, max(3, 0.5 / (B * 1.0e-9))))):
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
    value = G * args.steps / (ms * 1e-3)
    valid_fraction = None
    if strong:
        res = step()
        torch.cuda.synchronize()
        vloc = res[0][rank * B:(rank + 1) * B] if res[0].shape[0] == G and world > 1 else res[0]
        valid_fraction = float(vloc[: 1 << 22].float().mean().item())

    # ---- roofline of the dominant (only) kernel, per pipe: launch duration measured above (one launch per step)
    stats = ops.broadphase_stats(q[: 1 << 20], b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid, want_grad=True)
    cost = algorithmic_cost(W, robot, w["boxes"].shape[0], w["spheres"].shape[0], True, outputs,
                            pairs_per_state=stats["pairs_per_state"], lookups_per_state=stats["lookups_per_state"])
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
    # dense TF32 rate of this GPU (cuBLAS through torch.matmul, 8192^3, best of 5): the tensor-pipe denominator
    torch.backends.cuda.matmul.allow_tf32 = True
    a_ = torch.randn(8192, 8192, device=dev)
    b_ = torch.randn(8192, 8192, device=dev)
    (a_ @ b_)
    tt = []
    for _ in range(5):
        e0.record(); (a_ @ b_); e1.record(); torch.cuda.synchronize()
        tt.append(e0.elapsed_time(e1))
    tf32_peak = 2 * 8192 ** 3 / (min(tt) * 1e-3) / 1e12
    del a_, b_
    torch.backends.cuda.matmul.allow_tf32 = False
    step_s = ms * 1e-3 / args.steps
    rows_per_launch = B
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    tc_path = B > 16384
    per_step = launches / max(1, args.steps)
    two_pass = world == 1 and round(per_step) == 2 or (world > 1 and B >= ops.SPLIT_MIN_ROWS and os.environ.get("CBFRRT_SPLIT", "1") != "0")
    qp_flops = 2 * n * (3 * n + 4)

    def pipes_of(seconds, fp32, tensor, byts):
        d = {"fp32": {"achieved": fp32 * rows_per_launch / seconds / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                      "peak_source": "cbfrrt_fma_probe in this run (FP32 FMA rate; not in MEASURED_PEAKS.json)", "work_per_state": fp32},
             "tensor": {"achieved": tensor * rows_per_launch / seconds / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                        "peak_source": "torch.matmul TF32 8192^3 in this run (cuBLAS)", "work_per_state": tensor,
                        "note": "executed tcgen05 flops of the barrier network: 3-term TF32 split = 3 x the algorithmic 2*MACs"},
             "hbm": {"achieved": byts * rows_per_launch / seconds / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
                     "work_per_state": byts}}
        for p in d.values():
            p["frac"] = p["achieved"] / p["peak"]
        return d

    tr_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    tr_db = json.load(open(tr_file)) if os.path.exists(tr_file) else {}

    def traffic_of(key):
        ent = tr_db.get(key)
        if not ent:
            return None, "no ncu --set full capture committed for this kernel"
        return ent["dram_bytes_per_state"] * rows_per_launch, ent["source"]   # bytes per state of the capture x rows of THIS launch

    passes = None
    if two_pass:
        # the step is two launches: each pass is timed alone, live, with the same kernels / inputs / outputs (CUDA events on
        # the launching stream, L2 flushed in between); their sum against the step time is printed as a cross-check
        dx_s = torch.empty((B, 2 * n + 1), dtype=torch.float32, device=dev)
        safe_s, unsafe_s = torch.empty(B, dtype=torch.uint8, device=dev), torch.empty(B, dtype=torch.uint8, device=dev)
        u_s = torch.empty((B, n), dtype=torch.float32, device=dev)
        sc_c = ops._scene(b8, s4, 1, gr)

        def pass_a():
            ops.check(ops.lib.cbfrrt_collide(rid, ctypes.byref(sc_c), q.data_ptr(), n, B, w["thr_soft"], w["thr_hard"], safe_s.data_ptr(),
                                             None, dx_s.data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream), "cbfrrt_collide")

        def pass_b():
            ops.cbf_filter_u_into(dx_s, goal, net, K, lo, hi, w["lam"], w["penalty"], 48, 2, u_s)

        def timed_pass(fn):
            fn()
            tt_ = []
            for _ in range(3):
                flush.zero_()
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                tt_.append(e0.elapsed_time(e1) * 1e-3)
            return sum(tt_) / len(tt_)

        t_a, t_b = timed_pass(pass_a), timed_pass(pass_b)
        del dx_s, safe_s, unsafe_s, u_s
        in_b, dx_b = 4 * n, 4 * (2 * n + 1)
        passes = [
            {"kernel": f"collide_kernel<{robot}, grad, grid={bool(stats['grid'])}> (geometry pass: q -> valid, datax)", "ms": t_a * 1e3,
             "pipes": pipes_of(t_a, cost["fp32"] - qp_flops, 0, in_b + 1 + dx_b), "key": w["name"] + ":geometry_pass"},
            {"kernel": f"filter_kernel_tc<{n}> (barrier-network pass: datax -> u, tcgen05)", "ms": t_b * 1e3,
             "pipes": pipes_of(t_b, qp_flops, 3 * cost["mlp"], dx_b + in_b), "key": w["name"] + ":network_pass"},
        ]
        for p_ in passes:
            p_["bound"] = max(p_["pipes"], key=lambda k: p_["pipes"][k]["frac"])
            p_["frac"] = p_["pipes"][p_["bound"]]["frac"]
            p_["traffic"], p_["traffic_source"] = traffic_of(p_.pop("key"))
        dom = max(passes, key=lambda p_: p_["ms"])
        pipes, kernel_s, kernel_name = dom["pipes"], dom["ms"] * 1e-3, dom["kernel"]
        traffic, traffic_note = dom["traffic"], dom["traffic_source"]
    else:
        kernel_s = step_s
        pipes = pipes_of(kernel_s, cost["fp32"], 3 * cost["mlp"] if tc_path else 0, cost["bytes"])
        kernel_name = ("safety_kernel_tc" if tc_path else "safety_kernel") + f"<{robot}, grid={bool(stats['grid'])}>"
        traffic, traffic_note = traffic_of(w["name"])
    bound = max(pipes, key=lambda k: pipes[k]["frac"])
    roofline = {"bound": bound, "achieved": pipes[bound]["achieved"], "peak": pipes[bound]["peak"], "unit": pipes[bound]["unit"],
                "frac": pipes[bound]["frac"], "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                "traffic_source": traffic_note, "pipes": pipes, "kernel": kernel_name,
                "kernel_ms": kernel_s * 1e3, "rows_per_launch": rows_per_launch, "broad_phase": stats,
                "fp32_work_per_state_brute_force_formula": cost["fp32_brute_force"],
                "note": "per pipe, never summed: fp32 = algorithmic CUDA-core flops of the path as built (FK, centres, broad-"
                        "phase look-ups, the exact pair evaluations the broad phase leaves -- counted on the device for this "
                        "scene --, self pairs, gradient, QP) / kernel time against the measured FMA rate; tensor = executed "
                        "tcgen05 flops against a measured dense TF32 matmul; hbm = compulsory bytes against the measured "
                        "copy bandwidth.  frac = the largest.  SURVEY 8d's brute-force pair formula is kept beside it."}
    brute = cost["fp32_brute_force"]
    roofline["survey_8d_formula"] = {
        "fp32_flops_per_state": brute, "achieved": brute * rows_per_launch / step_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": brute * rows_per_launch / step_s / 1e12 / fp32_peak,
        "note": "SURVEY 8d's per-unit figure (every link sphere against every obstacle) x the units of one step / the step time: "
                "above 1 because the exact broad phase evaluates "
                f"{stats['pairs_per_state']:.1f} of the {W.ROBOT_MODELS[robot]['n_spheres'] if 'n_spheres' in W.ROBOT_MODELS[robot] else len(W.ROBOT_MODELS[robot]['spheres'])} x "
                f"{int(w['boxes'].shape[0] + w['spheres'].shape[0])} pairs the formula counts; the fractions above use the work actually left"}
    if passes is not None:
        roofline["passes"] = passes
        roofline["step_ms"] = step_s * 1e3
        roofline["passes_ms_sum"] = sum(p_["ms"] for p_ in passes)
        roofline["note"] += ("  The step is TWO launches (DESIGN.md 4.8): a geometry pass at its own occupancy and a tensor-core "
                             "barrier-network pass through the observation rows; the top-level figures are those of the longer "
                             "pass, `passes` holds both, each timed alone in this run.")

    # ---- end to end through the host-buffer C-ABI (pinned host buffers; H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        q_host = torch.empty((B, n), dtype=torch.float32, pin_memory=True)
        q_host.copy_(q)
        safe_h = torch.empty(B, dtype=torch.uint8, pin_memory=True)
        u_h = torch.empty((B, n), dtype=torch.float32, pin_memory=True)
        out = {"safe": safe_h.numpy(), "u": u_h.numpy()}
        q_host_np = q_host.numpy()
        def e2e_step():
            ops.safety_filter_host(robot, q_host_np, w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"],
                                   w["K"], w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"],
                                   out, device=local_rank)
        e2e_steps = max(3, min(args.steps, int(2.0 / max(B * 1.5e-9, 1e-4))))
        for _ in range(min(args.warmup, 3)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": G * e2e_steps / float(dt.item()), "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": int(B * n * 4), "d2h_bytes_per_step": int(B * (1 + 4 * n)),
               "api": "cbfrrt_safety_filter_host (C-ABI, pinned host q -> valid,u; bytes are per rank)"}
        del q_host, safe_h, u_h

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        qs = host_states(w, cpu_sample_rows(w))
        cpu_port_run(w, qs[:4096], threads)
        vals = [cpu_port_run(w, qs, threads) for _ in range(3)]
        cpu = {"value": float(np.median(vals)), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"the first {qs.shape[0]} of {G} states of {w['name']}, median of 3 (oracle/cbfrrt_oracle.c, pthreads)",
               "sequential": cpu_sequential_run(w, 256 if w["boxes"].shape[0] > 32 else 1024)}

    others = None
    if rank == 0 and world == 1 and strong and not args.no_other_configs and args.scale == 1.0:
        del q
        torch.cuda.empty_cache()
        others = other_configs(torch, ops, W, dev, flush)

    if rank == 0:
        targets = None
        if strong:
            targets = {k: {"target_8gpu": v, "value_at_this_n": value, "n_gpus": world, "met_at_this_n": bool(value >= v)}
                       for k, v in TARGETS_8GPU.items()}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, {
                "states": int(G), "states_per_gpu": int(B), "outputs": outputs, "valid_fraction": valid_fraction,
                "l2": "per-GPU inputs + outputs exceed the 126 MB L2; a 256 MiB write also flushes it between timed steps "
                      "(outside the event pairs)",
                "collective": collective, "host_cpus_bound": affinity}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": int(round(launches / max(1, args.steps))),
            "roofline": roofline, "cpu_baseline": cpu, "targets": targets, "other_configs": others,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
