"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Reads are sharded by id; every rank counts the k-mers of its shard.  The exchanges the path has:
(1) k-mers are owned by hash partition, so one all-to-all moves every (k-mer, count) pair to its owner, the owner
thresholds its share, and one all-gather hands every rank the union of survivors (`exchange_counts_and_finalize`);
(2) the reference-read set is global: one all-gather of every rank's reference reads, of which a rank keeps those of the
ranks before it as context reads (`exchange_reference_reads`), so that the candidates, tuples and streams of its shard are
the ones a single GPU would produce for the same reads.
`ctx` is a colord_b200.lib.Context (or, in the CPU gloo tests, any object with the same methods); all
buffers that cross NCCL are device tensors whose pointers go straight into the C-ABI.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

STAT_KEYS = ("n_reads", "tot_kmers", "n_unique", "n_unique_counted", "total_count_filtered")


class _Trace:
    """BENCH_PHASES=1: wall time of every step of an exchange on rank 0 (synchronising; debugging aid)."""

    def __init__(self, what, rank):
        import os
        import sys
        import time
        self.on = os.environ.get("BENCH_PHASES") is not None and rank == 0 and torch.cuda.is_available()
        self.what, self.time, self.err = what, time, sys.stderr
        if self.on:
            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def __call__(self, name):
        if self.on:
            torch.cuda.synchronize()
            t = self.time.perf_counter()
            print(f"[exchange] {self.what}: {name:22s} {1e3 * (t - self.t):8.1f} ms", file=self.err)
            self.t = t


def exchange_counts_and_finalize(ctx, device, n_local_reads: int, group=None):
    """Turn per-rank count tables into the identical global filtered set on every rank.

    Returns the global statistics dict.  World size 1 degenerates to ctx.count_finalize().
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ctx.count_finalize()
    rank = dist.get_rank(group)
    trace = _Trace("counts", rank)
    # 1. sizes of my table's partitions, exchanged so every rank knows what it will receive
    send_l = ctx.counts_sizes(world)                      # one pass over the table for all partitions
    send_sizes = torch.tensor(send_l, dtype=torch.int64, device=device)
    recv_sizes = torch.empty_like(send_sizes)
    dist.all_to_all_single(recv_sizes, send_sizes, group=group)
    recv_l = recv_sizes.tolist()
    trace("sizes")
    # 2. export all partitions into one send buffer (one more pass), all-to-all the pairs
    send_k = torch.empty(max(1, sum(send_l)), dtype=torch.int64, device=device)
    send_c = torch.empty(max(1, sum(send_l)), dtype=torch.int32, device=device)
    first = [sum(send_l[:p]) for p in range(world)]
    ctx.counts_export_all_device(world, first, send_k.data_ptr(), send_c.data_ptr(), max(1, sum(send_l)))
    trace("export")
    recv_k = torch.empty(max(1, sum(recv_l)), dtype=torch.int64, device=device)
    recv_c = torch.empty(max(1, sum(recv_l)), dtype=torch.int32, device=device)
    dist.all_to_all_single(recv_k[:sum(recv_l)], send_k[:sum(send_l)], recv_l, send_l, group=group)
    dist.all_to_all_single(recv_c[:sum(recv_l)], send_c[:sum(send_l)], recv_l, send_l, group=group)
    trace("all-to-all")
    # 3. my table now holds only what I own: everything every rank counted for my partition
    ctx.counts_reset()
    ctx.counts_merge_device(recv_k.data_ptr(), recv_c.data_ptr(), sum(recv_l), 0)
    trace("reset + merge")
    local = ctx.count_finalize()
    trace("finalize (local)")
    # 4. statistics are sums over owners (n_reads over shards)
    st = torch.tensor([n_local_reads] + [local[k] for k in STAT_KEYS[1:]], dtype=torch.int64, device=device)
    dist.all_reduce(st, group=group)
    stats = dict(zip(STAT_KEYS, st.tolist()))
    # 5. all-gather the survivors (padded to the largest share)
    n_mine = ctx.filter_size()
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    sizes[rank] = n_mine
    dist.all_reduce(sizes, group=group)
    sizes_l = sizes.tolist()
    pad = max(1, max(sizes_l))
    my_k = torch.zeros(pad, dtype=torch.int64, device=device)
    my_c = torch.zeros(pad, dtype=torch.int32, device=device)
    if n_mine:
        ctx.filter_list_device(my_k.data_ptr(), my_c.data_ptr(), n_mine)
    all_k = torch.empty(world * pad, dtype=torch.int64, device=device)
    all_c = torch.empty(world * pad, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(all_k, my_k, group=group)
    dist.all_gather_into_tensor(all_c, my_c, group=group)
    keep = torch.cat([torch.arange(r * pad, r * pad + sizes_l[r], device=device) for r in range(world)]) if sum(sizes_l) else torch.zeros(0, dtype=torch.int64, device=device)
    uni_k = all_k[keep].contiguous()
    uni_c = all_c[keep].contiguous()
    trace("survivors gathered")
    ctx.filter_import_device(uni_k.data_ptr(), uni_c.data_ptr(), int(uni_k.numel()), stats)
    trace("filter import")
    return stats


def exchange_reference_reads(ctx, device, sampled_local, lengths_local, group=None):
    """Make the reference-read set global.  Call after exchange_counts_and_finalize and before ctx.graph_build.

    sampled_local: the sampler's decisions for this rank's reads (numpy u8, the slice of the GLOBAL decision vector);
    lengths_local: their lengths (numpy).  Every rank exports its reference reads (sampled and free of N) as ASCII on the
    device; one all-gather (padded to the largest share) hands them to everybody; the reads of the ranks before this one
    become the context reads of `ctx`.  Returns the number of context reads (0 on rank 0 / world size 1).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return 0
    rank = dist.get_rank(group)
    n_local = len(sampled_local)
    has_n = ctx.reads_have_n(n_local)
    ids = np.nonzero((np.asarray(sampled_local) != 0) & (has_n == 0))[0].astype(np.uint32)
    lens = np.asarray(lengths_local)[ids].astype(np.int64)
    n_bases = int(lens.sum())
    sizes = torch.zeros(world, 2, dtype=torch.int64, device=device)
    sizes[rank, 0], sizes[rank, 1] = len(ids), n_bases
    dist.all_reduce(sizes, group=group)
    sizes_l = sizes.tolist()
    pad_r, pad_b = max(1, max(x[0] for x in sizes_l)), max(1, max(x[1] for x in sizes_l))
    my_lens = torch.zeros(pad_r, dtype=torch.int64, device=device)
    my_lens[:len(ids)] = torch.from_numpy(lens).to(device)
    my_bases = torch.zeros(pad_b, dtype=torch.uint8, device=device)
    if n_bases:
        ctx.reads_export(ids, n_bases, device_ptr=my_bases.data_ptr())
    all_lens = torch.empty(world * pad_r, dtype=torch.int64, device=device)
    all_bases = torch.empty(world * pad_b, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(all_lens, my_lens, group=group)
    dist.all_gather_into_tensor(all_bases, my_bases, group=group)
    n_ctx = sum(sizes_l[q][0] for q in range(rank))
    if n_ctx == 0:
        return 0
    ctx_lens = torch.cat([all_lens[q * pad_r:q * pad_r + sizes_l[q][0]] for q in range(rank)])
    ctx_bases = torch.cat([all_bases[q * pad_b:q * pad_b + sizes_l[q][1]] for q in range(rank)]).contiguous()
    ctx_off = torch.zeros(n_ctx + 1, dtype=torch.int64, device=device)
    ctx_off[1:] = torch.cumsum(ctx_lens, 0)
    del all_lens, all_bases
    ctx.append_context_reads(ctx_bases.data_ptr(), ctx_off.data_ptr(), n_ctx, on_device=True)
    return n_ctx
