"""One process per GPU: shard a state sweep across ranks, all-gather validity masks and filtered controls.

The path shards embarrassingly (independent states): scene, link model and MLP weights are
replicated, rows are partitioned, and the ONLY collective is the all-gather of ``valid`` (u8) and ``u`` (f32 x n)
that BASELINE.json's north_star names.  The reference has no counterpart (it fans out OS processes and merges
pickles offline, motion_planning/scripts/rrt_mindis.sh:8-15).

Partition: block-cyclic.  The G global rows are cut into blocks of ``block`` rows; block b belongs to rank
b % world.  A rank computes its blocks in order; block j of every rank is gathered by one
``all_gather_into_tensor`` whose output [world, block] is exactly the global rows
[j*world*block, (j+1)*world*block) -- so the gathered buffers are in global row order without a permute, and
the gather of block j overlaps the kernel of block j+1 on a second stream.
"""
import os
from typing import Callable, Tuple

import torch
import torch.distributed as dist


def block_layout(G: int, world: int, block: int):
    """-> (blocks_per_rank, padded global size).  G is padded up to a multiple of world*block."""
    per_round = world * block
    rounds = (G + per_round - 1) // per_round
    return rounds, rounds * per_round


def local_global_rows(rank: int, world: int, block: int, rounds: int) -> torch.Tensor:
    """global row index of every local row of ``rank`` (local rows are its blocks concatenated)."""
    j = torch.arange(rounds).repeat_interleave(block)
    i = torch.arange(block).repeat(rounds)
    return j * (world * block) + rank * block + i


class ShardedSweep:
    """valid, u = sweep(q_local): every rank ends up with the full [G] mask and [G, n] controls.

    compute(q_block) -> (valid u8[block], u f32[block, n]) is the per-block kernel (cbfrrt::valid_control on the
    GPU; tests inject a CPU function and the gloo backend).
    """

    def __init__(self, compute: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], n: int, block: int,
                 group=None, device=None):
        self.compute, self.n, self.block, self.group = compute, n, block, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._bufs = {}

    def _out(self, rounds):
        key = rounds
        if key not in self._bufs:
            G = rounds * self.world * self.block
            self._bufs[key] = (torch.empty(G, dtype=torch.uint8, device=self.device),
                               torch.empty(G, self.n, dtype=torch.float32, device=self.device))
        return self._bufs[key]

    def __call__(self, q_local: torch.Tensor):
        rounds = q_local.shape[0] // self.block
        assert rounds * self.block == q_local.shape[0], "local rows must be a whole number of blocks"
        valid_all, u_all = self._out(rounds)
        per_round = self.world * self.block
        for j in range(rounds):
            v, u = self.compute(q_local[j * self.block:(j + 1) * self.block])
            vo = valid_all[j * per_round:(j + 1) * per_round]
            uo = u_all[j * per_round:(j + 1) * per_round]
            if self.world == 1:
                vo.copy_(v)
                uo.copy_(u)
                continue
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    dist.all_gather_into_tensor(vo, v, group=self.group)
                    dist.all_gather_into_tensor(uo.view(-1), u.reshape(-1), group=self.group)
                    v.record_stream(self.comm_stream)
                    u.record_stream(self.comm_stream)
            else:
                dist.all_gather_into_tensor(vo, v.contiguous(), group=self.group)
                dist.all_gather_into_tensor(uo.view(-1), u.reshape(-1).contiguous(), group=self.group)
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        return valid_all, u_all


class FusedSweep:
    """Sweep + all-gather in ONE kernel per rank: the epilogue of the fused safety-filter kernel stores every
    (valid, u) row into all ranks' replicated buffers over NVLink (torch symmetric memory), then the ranks meet at a
    barrier.  Contiguous sharding: rank r owns global rows [r*B, (r+1)*B).

    On NVSwitch systems the stores go to the buffers' MULTICAST address: one store per row, replicated to every rank
    by the switch (NVLS) -- 1/world of the outbound traffic and store instructions of the per-peer form, which is
    kept as the fallback (no multicast support, or CBFRRT_NO_MULTICAST=1).

    launch(q_local, valid_ptrs, u_ptrs, row_offset, multicast) runs cbfrrt::valid_control_scatter with the scene /
    network of the caller.  ``valid`` [G] u8 and ``u`` [G, n] f32 are this rank's replicas (valid after __call__
    returns and the stream is synchronised).

    Ordering contract: a step begins with a barrier too -- no rank's kernel may store into the replicas while another
    rank still reads the previous step's rows (write-after-read across ranks).  Both barriers are stream-ordered
    symmetric-memory barriers: whatever the caller enqueued on the current stream to consume step i is complete on
    every rank before any store of step i+1 is issued.  The multicast form needs rows_per_rank % 128 == 0 (the mask
    bytes of 32 rows travel as packed words); otherwise the per-peer form is used."""

    # the step is exchange-bound when the stores of a step (every rank receives G rows of 4 n + 1 bytes over NVLink,
    # ~800 GB/s) take longer than the kernel pass that issues them (the barrier-network pass of the two-pass sweep,
    # 0.17 ns per local row on B200): 8 GPUs on 2^26 Panda states (2.4 ms of stores behind a 1.4 ms pass).  Measured
    # there (ms per step): one chunk 6.20, 2 / 4 / 8 / 16 chunks 6.01 / 5.81 / 5.58 / 5.16, the single fused kernel
    # 5.86 (profiles/r2y_scale_n8_*).  Chunks of 2^19 rows keep the tail of each launch small.
    NVLINK_BYTES_PER_S, NETWORK_PASS_S_PER_ROW, MIN_CHUNK_ROWS, MAX_CHUNKS = 800e9, 0.17e-9, 1 << 19, 16

    @classmethod
    def choose_chunks(cls, n: int, rows_per_rank: int, world: int) -> int:
        """1 unless the step is exchange-bound; then the largest power of two <= MAX_CHUNKS whose chunks keep MIN_CHUNK_ROWS
        rows and whole 128-row tiles"""
        bound = world * rows_per_rank * (4 * n + 1) / cls.NVLINK_BYTES_PER_S > rows_per_rank * cls.NETWORK_PASS_S_PER_ROW
        chunks = 1
        while bound and chunks < cls.MAX_CHUNKS and rows_per_rank // (2 * chunks) >= cls.MIN_CHUNK_ROWS \
                and rows_per_rank % (256 * chunks) == 0:
            chunks *= 2
        return chunks

    def __init__(self, launch, n: int, rows_per_rank: int, device, group=None, passes=None, chunks=None):
        """passes = (geometry(q_rows, safe_out, datax_out), network(datax, safe, rows, valid_ptrs, u_ptrs, row_offset,
        multicast)): the two passes of the sweep as separate calls (ops.observe_into / ops.cbf_filter_scatter).  Given
        them, an exchange-bound step (or chunks > 1) is PIPELINED in chunks over two streams: the geometry pass of chunk
        k + 1 runs while the NVLink stores issued by the barrier-network pass of chunk k drain (a kernel's CTAs retire
        long before its posted remote stores are acknowledged; in one stream the next kernel would wait for them)."""
        import torch.distributed._symmetric_memory as symm_mem
        self.launch, self.n, self.B = launch, n, rows_per_rank
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        G = self.world * rows_per_rank
        self.valid = symm_mem.empty(G, dtype=torch.uint8, device=device)
        self.u = symm_mem.empty(G, n, dtype=torch.float32, device=device)
        self.h_valid = symm_mem.rendezvous(self.valid, self.group)
        self.h_u = symm_mem.rendezvous(self.u, self.group)
        self.valid_ptrs = [int(p) for p in self.h_valid.buffer_ptrs]
        self.u_ptrs = [int(p) for p in self.h_u.buffer_ptrs]
        mc_v = int(getattr(self.h_valid, "multicast_ptr", 0) or 0)
        mc_u = int(getattr(self.h_u, "multicast_ptr", 0) or 0)
        self.multicast = (bool(mc_v and mc_u) and os.environ.get("CBFRRT_NO_MULTICAST", "0") != "1"
                          and rows_per_rank % 128 == 0)
        if self.multicast:
            self.valid_ptrs, self.u_ptrs = [mc_v], [mc_u]
        self.passes, self.chunks = passes, 1
        if passes is not None:
            if chunks is None:
                chunks = self.choose_chunks(n, rows_per_rank, self.world)
            if chunks > 1 and rows_per_rank % (128 * chunks) == 0:
                self.chunks = int(chunks)
                bc = rows_per_rank // self.chunks
                self.side = torch.cuda.Stream(device=device, priority=-1)
                self.scratch = [(torch.empty(bc, dtype=torch.uint8, device=device),
                                 torch.empty((bc, 2 * n + 1), dtype=torch.float32, device=device)) for _ in range(3)]

    def __call__(self, q_local: torch.Tensor):
        assert q_local.shape[0] == self.B
        self.h_u.barrier(channel=1)     # every rank is done reading the previous step's rows
        if self.chunks == 1:
            self.launch(q_local, self.valid_ptrs, self.u_ptrs, self.rank * self.B, self.multicast)
        else:
            geometry, network = self.passes
            cur, bc, done = torch.cuda.current_stream(), self.B // self.chunks, []
            for c in range(self.chunks):
                safe_c, datax_c = self.scratch[c % 3]
                if c >= 3:
                    cur.wait_event(done[c - 3])          # the network pass that read this scratch has finished
                geometry(q_local[c * bc:(c + 1) * bc], safe_c, datax_c)
                ready = cur.record_event()
                with torch.cuda.stream(self.side):
                    self.side.wait_event(ready)
                    network(datax_c, safe_c, bc, self.valid_ptrs, self.u_ptrs, self.rank * self.B + c * bc, self.multicast)
                    done.append(self.side.record_event())
            for ev in done[-3:]:
                cur.wait_event(ev)
        self.h_u.barrier(channel=0)     # all ranks' kernels (and their peer stores) are complete past this point
        return self.valid, self.u
