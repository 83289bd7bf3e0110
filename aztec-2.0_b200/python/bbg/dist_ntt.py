"""Multi-GPU four-step NTT plumbing (one process per GPU, torch.distributed over NCCL).

The device work is libbbg's (bbg_ntt_dist_dev, csrc/ntt.cu): phase 0 = every pass but the last on this rank's packed
sub-array, phase 1 = the last pass after ONE all-to-all of equal contiguous chunks.  This module only moves chunks
(torch.distributed.all_to_all_single) and converts between a natural-order array and the per-rank layouts:

    input  shard of rank r : x[i] with bits [in_pos,  in_pos  + log2 world) of i == r, packed in index order
    output shard of rank r : X[k] with bits [out_pos, out_pos + log2 world) of k == r, packed in index order

`simulate` runs all ranks one after the other on ONE device with the exchange done by slicing -- the same kernels and
index maths as the multi-process path -- so the N > 1 data path is parity-tested on a single GPU.
"""
import numpy as np


def extract_shard(x, pos, world, rank):
    """x: (n, 4) natural order (numpy or torch). Returns the packed (n / world, 4) sub-array of `rank`."""
    n = x.shape[0]
    low = 1 << pos
    return x.reshape(n // (low * world), world, low, 4)[:, rank].reshape(n // world, 4)


def insert_shard(out, shard, pos, world, rank):
    n = out.shape[0]
    low = 1 << pos
    out.reshape(n // (low * world), world, low, 4)[:, rank] = shard.reshape(n // (low * world), low, 4)


def ntt_sharded(bbg, local_in, n, kind, rank, world, generator_size=0, constant=None, group=None, phase_fn=None):
    """local_in: torch int64 tensor (n / world, 4) = this rank's input shard (CUDA in the product).  Returns this rank's
    output shard.  phase_fn(src, dst, n, kind, rank, world, phase, generator_size, constant) defaults to libbbg's
    bbg_ntt_dist_dev; tests/test_sharded_cpu.py injects a CPU stand-in to run this plumbing under gloo."""
    import torch
    import torch.distributed as dist
    phase = bbg.ntt_dist_phase if phase_fn is None else phase_fn
    mid = torch.empty_like(local_in)
    phase(local_in, mid, n, kind, rank, world, 0, generator_size, constant)
    if world == 1:
        return mid
    recv = torch.empty_like(mid)
    dist.all_to_all_single(recv, mid, group=group)  # chunk r of every rank -> rank r, in source-rank order
    out = torch.empty_like(mid)
    phase(recv, out, n, kind, rank, world, 1, generator_size, constant)
    return out


class FusedExchange:
    """Peer-mapped receive buffers for ntt_sharded_fused: two per rank (alternated between consecutive transforms so
    that a peer's stores for transform k + 1 never land in the buffer this rank is still reading for transform k),
    IPC handles exchanged once through torch.distributed."""

    def __init__(self, bbg, n, rank, world, group=None, natural=False):
        import torch
        import torch.distributed as dist
        self.bbg, self.n, self.rank, self.world, self.group = bbg, n, rank, world, group
        self.m = n // world
        self.own, self.peers = [], []
        # buffers 0, 1: receive buffers; natural: buffer 2 = this rank's natural input block, 3 = its natural output block
        for _ in range(4 if natural else 2):
            ptr, handle = bbg.peer_buffer_alloc(self.m * 32)
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            ptrs = [ptr if r == rank else bbg.peer_buffer_open(handles[r]) for r in range(world)]
            self.own.append(ptr)
            self.peers.append(ptrs)
        if natural:
            self.in_view = torch.as_tensor(bbg.RawCudaArray(self.own[2], self.m), device="cuda")
            self.out_view = torch.as_tensor(bbg.RawCudaArray(self.own[3], self.m), device="cuda")
        self.turn = 0
        self.token = torch.zeros(1, dtype=torch.float32, device="cuda")
        dist.barrier(group=group)

    def close(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.in_view = self.out_view = None
        for b in range(len(self.own)):
            for r in range(self.world):
                if r != self.rank:
                    self.bbg.peer_buffer_close(self.peers[b][r])
        dist.barrier(group=self.group)
        for b in range(len(self.own)):
            self.bbg.peer_buffer_free(self.own[b])


def ntt_sharded_fused(bbg, local_in, n, kind, rank, world, xch, generator_size=0, constant=None):
    """ntt_sharded with the all-to-all FUSED into the pass before it: that pass stores its results straight into the
    owners' receive buffers over NVLink peer memory (bbg_ntt_dist_fused_dev), so the transfer overlaps the pass's
    arithmetic and no NCCL all-to-all runs.  The only collective left is a one-word all-reduce that orders phase 1 after
    every rank's stores (stream-ordered; the host does not wait).  Same layouts and bit-exact results as ntt_sharded."""
    import torch
    import torch.distributed as dist
    b = xch.turn
    xch.turn ^= 1
    work = torch.empty_like(local_in)
    bbg.ntt_dist_fused_phase0(local_in, work, xch.peers[b], n, kind, rank, world, generator_size, constant)
    dist.all_reduce(xch.token, group=xch.group)  # every rank's peer stores are complete when this completes
    out = torch.empty_like(local_in)
    bbg.ntt_dist_phase_raw(xch.own[b], out, n, kind, rank, world, 1, generator_size, constant)
    return out


def ntt_natural_fused(bbg, block_in, n, kind, rank, world, xch, generator_size=0, constant=None):
    """SURVEY.md 8e contract -- natural contiguous block in, natural block out -- with NO all-to-all and no re-distribution:
    every movement is peer-memory traffic issued by the NTT passes themselves (bbg_ntt_dist_natural_dev): the first pass
    loads its sliced sub-array from the owners' input blocks, the pass before the transposition stores into the owners'
    receive buffers, the last pass stores every output to the owner of its natural index.  Three one-word all-reduces order
    the phases across ranks.  xch: FusedExchange(..., natural=True); block_in may be xch.in_view itself (no copy).  Returns
    xch.out_view, valid until the next transform through the same exchange."""
    import torch.distributed as dist
    if block_in.data_ptr() != xch.in_view.data_ptr():
        xch.in_view.copy_(block_in)
    b = xch.turn
    xch.turn ^= 1
    work = xch.work if getattr(xch, "work", None) is not None else None
    if work is None:
        import torch
        work = xch.work = torch.empty_like(xch.in_view)
    dist.all_reduce(xch.token, group=xch.group)  # every rank's input block is in place
    bbg.ntt_dist_natural_phase(xch.peers[2], work, xch.peers[b], None, n, kind, rank, world, 0, generator_size, constant)
    dist.all_reduce(xch.token, group=xch.group)  # every rank's stores into the receive buffers are complete
    bbg.ntt_dist_natural_phase(None, None, xch.peers[b], xch.peers[3], n, kind, rank, world, 1, generator_size, constant)
    dist.all_reduce(xch.token, group=xch.group)  # every output block is complete
    return xch.out_view


def ntt_sharded_natural(bbg, block_in, n, kind, rank, world, generator_size=0, constant=None, group=None, phase_fn=None, layout=None):
    """SURVEY.md 8e contract: rank r holds the NATURAL contiguous block x[r n/W, (r+1) n/W) and ends with the natural
    block X[r n/W, (r+1) n/W).  The four-step core wants its input sliced by index bits [in_pos, in_pos + k) and leaves its
    output sliced by bits [out_pos, ...); the two re-distributions are one all-to-all each:
      in : element i of my block goes to rank bits(i); packed by (destination, index) the chunks received in rank order
           ARE the destination's packed shard (global index order = source block major);
      out: my packed output shard is ordered by global index, so the part belonging to rank q's block is the contiguous
           chunk q; the receiver interleaves the chunks by the slicing bits.
    Three all-to-alls in total (the sliced-layout entry point `ntt_sharded` needs one and remains the fast path when the
    caller can produce / consume the sliced layout)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return ntt_sharded(bbg, block_in, n, kind, rank, world, generator_size, constant, group, phase_fn)
    in_pos, out_pos = layout if layout is not None else bbg.ntt_dist_layout(n, world)
    m = n // world
    lb = m.bit_length() - 1  # bits of a block-local index; the slicing bits lie below the block bits (in_pos + k <= lb)
    assert in_pos + (world.bit_length() - 1) <= lb and out_pos + (world.bit_length() - 1) <= lb
    low = 1 << in_pos
    # pack: (m / (low W), W, low, 4) -> destination-major
    send = block_in.reshape(m // (low * world), world, low, 4).permute(1, 0, 2, 3).contiguous().reshape(m, 4)
    shard_in = torch.empty_like(send)
    dist.all_to_all_single(shard_in, send, group=group)
    shard_out = ntt_sharded(bbg, shard_in, n, kind, rank, world, generator_size, constant, group, phase_fn)
    recv = torch.empty_like(shard_out)
    dist.all_to_all_single(recv, shard_out.contiguous(), group=group)  # chunk q of my shard = my part of rank q's block
    low = 1 << out_pos
    # unpack: chunk s holds the elements of my block whose slicing bits equal s, in index order
    block_out = recv.reshape(world, m // (low * world), low, 4).permute(1, 0, 2, 3).contiguous().reshape(m, 4)
    return block_out


def simulate(bbg, x, kind, world, generator_size=0, constant=None, fused=False):
    """All `world` ranks on one device. x: torch CUDA int64 (n, 4) natural order -> natural-order result.
    fused: the exchange is done by the pass before it storing into the "peers'" receive buffers (here: buffers of the
    same device) -- the kernel and index maths of bbg_ntt_dist_fused_dev.  fused = "natural": natural blocks in and out
    through peer loads / stores as well (bbg_ntt_dist_natural_dev)."""
    import torch
    n = x.shape[0]
    in_pos, out_pos = bbg.ntt_dist_layout(n, world)
    m = n // world
    if fused == "natural":
        ins = [x[r * m:(r + 1) * m].contiguous() for r in range(world)]
        recvs = [torch.zeros((m, 4), dtype=x.dtype, device=x.device) for _ in range(world)]
        outs = [torch.zeros((m, 4), dtype=x.dtype, device=x.device) for _ in range(world)]
        pin, prc, pout = ([t.data_ptr() for t in ts] for ts in (ins, recvs, outs))
        for r in range(world):
            bbg.ntt_dist_natural_phase(pin, torch.empty_like(ins[r]), prc, None, n, kind, r, world, 0, generator_size, constant)
        for r in range(world):
            bbg.ntt_dist_natural_phase(None, None, prc, pout, n, kind, r, world, 1, generator_size, constant)
        torch.cuda.synchronize()
        return torch.cat(outs, dim=0)
    if fused:
        recvs = [torch.zeros((m, 4), dtype=x.dtype, device=x.device) for _ in range(world)]
        ptrs = [t.data_ptr() for t in recvs]
        for r in range(world):
            src = extract_shard(x, in_pos, world, r).contiguous()
            bbg.ntt_dist_fused_phase0(src, torch.empty_like(src), ptrs, n, kind, r, world, generator_size, constant)
        out = torch.empty_like(x)
        for r in range(world):
            dst = torch.empty_like(recvs[r])
            bbg.ntt_dist_phase(recvs[r], dst, n, kind, r, world, 1, generator_size, constant)
            insert_shard(out, dst, out_pos, world, r)
        torch.cuda.synchronize()
        return out
    mids = []
    for r in range(world):
        src = extract_shard(x, in_pos, world, r).contiguous()
        mid = torch.empty_like(src)
        bbg.ntt_dist_phase(src, mid, n, kind, r, world, 0, generator_size, constant)
        mids.append(mid)
    out = torch.empty_like(x)
    c = m // world
    for r in range(world):
        recv = torch.cat([mids[s][r * c:(r + 1) * c] for s in range(world)], dim=0).contiguous()
        dst = torch.empty_like(recv)
        bbg.ntt_dist_phase(recv, dst, n, kind, r, world, 1, generator_size, constant)
        insert_shard(out, dst, out_pos, world, r)
    torch.cuda.synchronize()
    return out
