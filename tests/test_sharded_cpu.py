"""CPU suite: the N > 1 path (block-cyclic sharding + all-gather of masks and controls) on world_size 2 with the
gloo backend; the per-block compute is the oracle-backed fake of cbfrrt::valid_control."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, block, rounds, ret):
    for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cbfrrt import workloads as W
    from cbfrrt.sharded import ShardedSweep, local_global_rows
    from oracle import c_oracle
    w = W.config(0, scale=0.01)
    G = world * block * rounds
    q_global = c_oracle.sample_states("panda", 3, 0, G)                      # counter-based: any rank can make any row
    rows = local_global_rows(rank, world, block, rounds).numpy()
    q_local = torch.from_numpy(q_global[rows])

    def compute(qb):
        o = c_oracle.safety_filter("panda", qb.numpy(), w["boxes"], None, True, w["weights"], 48, 2, w["goal"], w["K"],
                                   w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
        return torch.from_numpy(o["safe"]), torch.from_numpy(o["u"])

    valid, u = ShardedSweep(compute, 7, block, device="cpu")(q_local)
    ref_v, ref_u = compute(torch.from_numpy(q_global))                          # single-process answer, global order
    ok = torch.equal(valid, ref_v) and torch.equal(u, ref_u)
    gathered = [None] * world
    dist.all_gather_object(gathered, (bool(ok), int(valid.sum())))
    if rank == 0:
        ret.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("block,rounds", [(64, 3), (7, 1)])
def test_block_cyclic_sweep_world2_gloo(coracle, block, rounds):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + block
    procs = [ctx.Process(target=_worker, args=(r, 2, port, block, rounds, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in out) and out[0][1] == out[1][1]


def test_layout_helpers():
    from cbfrrt.sharded import block_layout, local_global_rows
    rounds, G = block_layout(1000, 4, 64)
    assert rounds == 4 and G == 1024
    rows = torch.cat([local_global_rows(r, 4, 64, rounds) for r in range(4)])
    assert torch.equal(rows.sort().values, torch.arange(G))
    assert local_global_rows(1, 4, 64, rounds)[:2].tolist() == [64, 65] and local_global_rows(1, 4, 64, rounds)[64].item() == 320


def test_fused_sweep_chunk_policy():
    """FusedSweep pipelines a step only where the exchange bounds it (every rank receives all G rows of 4 n + 1 bytes per
    step, whatever the number of GPUs): BASELINE config 5 (2^26 Panda states) at 8 GPUs, not at 2 or 4; chunks keep whole
    128-row tiles and at least 2^19 rows"""
    from cbfrrt.sharded import FusedSweep
    G = 1 << 26
    assert [FusedSweep.choose_chunks(7, G // w, w) for w in (2, 4, 8)] == [1, 1, 16]
    assert FusedSweep.choose_chunks(4, (1 << 20) // 8, 8) == 1                  # too few rows per rank to cut
    assert FusedSweep.choose_chunks(7, (1 << 23) + 128, 8) == 1                 # rows per rank not a multiple of 256
    assert FusedSweep.choose_chunks(7, 3 << 21, 8) == 8                         # 6 291 456 rows: 8 chunks of 786 432
    for rows, w in ((1 << 23, 8), (3 << 21, 8), (1 << 22, 16)):
        c = FusedSweep.choose_chunks(7, rows, w)
        assert rows % (128 * c) == 0 and (c == 1 or rows // c >= FusedSweep.MIN_CHUNK_ROWS)
