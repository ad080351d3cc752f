"""GPU suite, >= 2 GPUs: the fused sweep + all-gather (kernel epilogue stores into every rank's replica over NVLink
peer mappings) and the NCCL block-cyclic path both reproduce the single-GPU result bit for bit."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from cbfrrt import ops, workloads as W
    from cbfrrt.sharded import FusedSweep, ShardedSweep, local_global_rows
    w = W.config(1, scale=1 / 16)
    n, rid, B = w["n"], 1, 20096      # a multiple of 128: the multicast form packs the mask bytes of 32-row groups into words
    G = world * B
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
    b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
    gr = ops.build_grid(b8, s4)
    net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
    common = (t(w["goal"]), b8, s4, 1, gr, net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), w["lam"], w["penalty"],
              w["thr_soft"], w["thr_hard"], rid, 48, 2)
    q_all = ops.sample_states(G, 5, 0, rid, device=dev)              # counter-based: identical on every rank
    ref_v, ref_u = ops.valid_control(q_all, *common)
    ok = []
    # fused epilogue exchange, contiguous shards: multicast stores (NVSwitch) where available, then per-peer stores
    modes = []
    for no_mc in ("0", "1"):
        os.environ["CBFRRT_NO_MULTICAST"] = no_mc
        sweep = FusedSweep(lambda qb, vp, up, off, mc: ops.valid_control_scatter(qb, *common, vp, up, off, mc), n, B, dev)
        modes.append(sweep.multicast)
        for _ in range(2):
            sweep.valid.zero_()
            sweep.u.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            v, u = sweep(q_all[rank * B:(rank + 1) * B])
            torch.cuda.synchronize()
            ok.append(bool(torch.equal(v, ref_v) and torch.equal(u, ref_u)))
        dist.barrier()
    assert modes[1] is False
    # rows per rank not a multiple of 128: the sweep must fall back to per-peer stores by itself and still be exact
    os.environ["CBFRRT_NO_MULTICAST"] = "0"
    B2 = 19999
    q2 = ops.sample_states(world * B2, 6, 0, rid, device=dev)
    ref_v2, ref_u2 = ops.valid_control(q2, *common)
    sweep2 = FusedSweep(lambda qb, vp, up, off, mc: ops.valid_control_scatter(qb, *common, vp, up, off, mc), n, B2, dev)
    assert sweep2.multicast is False
    v, u = sweep2(q2[rank * B2:(rank + 1) * B2])
    torch.cuda.synchronize()
    ok.append(bool(torch.equal(v, ref_v2) and torch.equal(u, ref_u2)))
    dist.barrier()
    # >= 2^19 rows per rank: the two-pass sweep, the peer stores (multicast, then per-peer) issued by its second pass
    B3 = (1 << 19) + 128
    q3 = ops.sample_states(world * B3, 7, 0, rid, device=dev)
    ops.set_two_pass(False)
    ref_v3, ref_u3 = ops.valid_control(q3, *common)                   # single fused kernel
    ops.set_two_pass(True)
    for no_mc in ("0", "1"):
        os.environ["CBFRRT_NO_MULTICAST"] = no_mc
        sweep3 = FusedSweep(lambda qb, vp, up, off, mc: ops.valid_control_scatter(qb, *common, vp, up, off, mc), n, B3, dev)
        n0 = ops.launch_count()
        v, u = sweep3(q3[rank * B3:(rank + 1) * B3])
        torch.cuda.synchronize()
        ok.append(bool(torch.equal(v, ref_v3) and torch.equal(u, ref_u3)) and ops.launch_count() - n0 == 2)
        dist.barrier()
        del sweep3
    os.environ["CBFRRT_NO_MULTICAST"] = "0"
    # the same sweep pipelined in chunks over two streams (geometry pass / barrier-network pass with the peer stores)
    B4 = (1 << 18) + 512
    q4 = ops.sample_states(world * B4, 8, 0, rid, device=dev)
    ref_v4, ref_u4 = ops.valid_control(q4, *common)
    goal_t, K_t, lo_t, hi_t = common[0], common[6], common[7], common[8]
    passes = (lambda qb, s_out, d_out: ops.observe_into(qb, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid, s_out, d_out),
              lambda dx, sv, rows, vp, up, off, mc: ops.cbf_filter_scatter(dx, sv, rows, goal_t, net, K_t, lo_t, hi_t, w["lam"], w["penalty"],
                                                                          48, 2, n, vp, up, off, mc))
    for chunks in (4, 2):
        sweep4 = FusedSweep(lambda qb, vp, up, off, mc: ops.valid_control_scatter(qb, *common, vp, up, off, mc), n, B4, dev,
                            passes=passes, chunks=chunks)
        assert sweep4.chunks == chunks
        for _ in range(2):
            v, u = sweep4(q4[rank * B4:(rank + 1) * B4])
            torch.cuda.synchronize()
            ok.append(bool(torch.equal(v, ref_v4) and torch.equal(u, ref_u4)))
        dist.barrier()
        del sweep4
    if rank == 0:
        print("fused exchange modes tested (multicast, per-peer):", modes, flush=True)
    # NCCL all-gather, block-cyclic
    block, rounds = 2512, B // 2512
    rows = local_global_rows(rank, world, block, rounds).to(dev)
    v, u = ShardedSweep(lambda qb: ops.valid_control(qb, *common), n, block, device=dev)(q_all[rows])
    torch.cuda.synchronize()
    # blocks of 2500 rows take the FP32-FMA kernels (small-batch rule), the 40000-row reference the tcgen05 ones:
    # masks are bit-exact, controls agree to the 1e-5 tolerance of the two MLP paths
    err = ((u - ref_u).abs() / ref_u.abs().clamp(min=1.0)).max().item()
    ok.append(bool(torch.equal(v, ref_v) and err < 1e-3 and ((u - ref_u).abs() / ref_u.abs().clamp(min=1.0) < 1e-5).float().mean().item() > 0.995))
    out = [None] * world
    dist.all_gather_object(out, ok)
    if rank == 0:
        ret.put(out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_and_nccl_exchange_match_single_gpu():
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29620 + os.getpid() % 300, ret)) for r in range(world)]
    for p in procs:
        p.start()
    out = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(all(o) for o in out), out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_host_api_two_devices_one_process():
    """cbfrrt_safety_filter_host driving two GPUs from ONE process: the > 48 KB dynamic shared-memory opt-in and the SM
    count are per device (a cache keyed by kernel only would fail the second device's first large launch), and the
    caller's current device is left as it was"""
    import numpy as np
    for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from cbfrrt import ops, workloads as W
    w = W.config(1, scale=1 / 16)
    B, n = w["q"].shape
    outs = []
    torch.cuda.set_device(0)
    for dev in (0, 1, 0):
        out = dict(safe=np.empty(B, np.uint8), u=np.empty((B, n), np.float32), h=np.empty(B, np.float32))
        ops.safety_filter_host("magician", w["q"], w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"], w["K"],
                               w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], out, device=dev)
        assert torch.cuda.current_device() == 0
        outs.append(out)
    for k in ("safe", "u", "h"):
        assert np.array_equal(outs[0][k], outs[1][k]) and np.array_equal(outs[0][k], outs[2][k])
    # the device-pointer ABI on the second device as well (tcgen05 kernel, ~180 KB of dynamic shared memory)
    res = []
    for d in (0, 1):
        dv = torch.device("cuda", d)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dv)
        b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dv)
        gr = ops.build_grid(b8, s4)
        net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dv)
        v, u = ops.valid_control(t(w["q"]), t(w["goal"]), b8, s4, 1, gr, net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), w["lam"],
                                 w["penalty"], w["thr_soft"], w["thr_hard"], 1, 48, 2)
        res.append((v.cpu(), u.cpu()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][0].numpy(), outs[0]["safe"]) and np.array_equal(res[0][1].numpy(), outs[0]["u"])
