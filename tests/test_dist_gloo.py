"""Host-side logic of the multi-GPU exchanges (colord_b200/dist.py) under torch.distributed / gloo with two processes on CPU.

The device context is replaced by a stand-in with the same methods, backed by the C oracle (tests may use it): what is checked is
the plumbing — partition ownership, the all-to-all of (k-mer, count) pairs, thresholding by the owner, the all-gather of the
survivors, the summed statistics, and the all-gather of the reference reads into each rank's context reads — against the
single-process result on the whole input."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib
from colord_b200 import synth
from colord_b200.dist import exchange_counts_and_finalize, exchange_reference_reads

K, F, LO, HI = 20, 9, 3, 100
N_READS, GENOME, MEAN = 240, 60000, 1500


def _view(ptr, n, dtype):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(max(n, 1),))[:n]


class OracleCtx:
    """The methods of colord_b200.lib.Context that dist.py calls, over host memory."""

    def __init__(self, bases, offsets):
        self.bases, self.offsets = bases, offsets
        km, ct, st = oracle_lib.count_kmers(bases, offsets, K, F, 1, 1 << 30)      # every passing k-mer with its true count
        self.table = dict(zip(km.tolist(), ct.tolist()))
        self.tot_kmers = st["tot_kmers"]
        self.context = None

    @staticmethod
    def owner(kmers, n_parts):
        return (kmers * np.uint64(0x9E3779B97F4A7C15) >> np.uint64(40)) % np.uint64(n_parts)

    def _part(self, part, n_parts):
        km = np.fromiter(self.table.keys(), np.uint64, len(self.table))
        return km[self.owner(km, n_parts) == part]

    def counts_size(self, part, n_parts):
        return len(self._part(part, n_parts))

    def counts_export_device(self, part, n_parts, kptr, cptr, cap):
        km = self._part(part, n_parts)
        _view(kptr, len(km), np.uint64)[:] = km
        _view(cptr, len(km), np.uint32)[:] = [self.table[k] for k in km.tolist()]
        return len(km)

    def counts_sizes(self, n_parts):
        return [self.counts_size(p, n_parts) for p in range(n_parts)]

    def counts_export_all_device(self, n_parts, first, kptr, cptr, cap):
        for p in range(n_parts):
            km = self._part(p, n_parts)
            _view(kptr, cap, np.uint64)[first[p]:first[p] + len(km)] = km
            _view(cptr, cap, np.uint32)[first[p]:first[p] + len(km)] = [self.table[k] for k in km.tolist()]

    def counts_reset(self):
        self.table = {}

    def counts_merge_device(self, kptr, cptr, n, n_reads_remote=0):
        for k, c in zip(_view(kptr, n, np.uint64).tolist(), _view(cptr, n, np.uint32).tolist()):
            self.table[k] = self.table.get(k, 0) + c

    def count_finalize(self):      # kb_sorter.h:1011-1065: keep count >= L, saturate at H
        tot = sum(self.table.values())
        self.filtered = {k: min(c, HI) for k, c in self.table.items() if c >= LO}
        return dict(n_reads=len(self.offsets) - 1, tot_kmers=tot, n_unique=len(self.table), n_unique_counted=len(self.filtered),
                    total_count_filtered=sum(self.filtered.values()))

    def filter_size(self):
        return len(self.filtered)

    def filter_list_device(self, kptr, cptr, n):
        _view(kptr, n, np.uint64)[:] = list(self.filtered.keys())
        _view(cptr, n, np.uint32)[:] = list(self.filtered.values())

    def filter_import_device(self, kptr, cptr, n, stats=None):
        self.filtered = dict(zip(_view(kptr, n, np.uint64).tolist(), _view(cptr, n, np.uint32).tolist()))
        self.stats = stats

    def reads_have_n(self, n):
        return np.array([(self.bases[int(self.offsets[i]):int(self.offsets[i + 1])] == ord("N")).any() for i in range(n)], np.uint8)

    def reads_export(self, ids, total, device_ptr=None):
        out = np.concatenate([self.bases[int(self.offsets[i]):int(self.offsets[i + 1])] for i in ids]) if len(ids) else np.zeros(0, np.uint8)
        assert len(out) == total
        _view(device_ptr, total, np.uint8)[:] = out

    def append_context_reads(self, bptr, optr, n, on_device=False):
        off = _view(optr, n + 1, np.int64).copy()
        self.context = (_view(bptr, int(off[-1]), np.uint8).copy(), off)


def _worker(rank, world, port, seed):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = synth.generate(N_READS, GENOME, MEAN, seed=seed, profile="ont", n_frac=0.05)
        lo, hi = rank * N_READS // world, (rank + 1) * N_READS // world
        o = s.offsets
        bases = s.bases[int(o[lo]):int(o[hi])]
        offsets = (o[lo:hi + 1] - o[lo]).astype(np.uint64)
        ctx = OracleCtx(bases, offsets)
        stats = exchange_counts_and_finalize(ctx, torch.device("cpu"), hi - lo)
        # the single-process answer on the whole input
        km, ct, st = oracle_lib.count_kmers(s.bases, s.offsets, K, F, LO, HI)
        assert stats == st, (stats, st)
        assert dict(zip(km.tolist(), ct.tolist())) == ctx.filtered
        # reference reads: every rank receives exactly the sampled N-free reads of the ranks before it, in input order
        sampled = oracle_lib.sampler(25, 1.0, 0, N_READS)
        n_ctx = exchange_reference_reads(ctx, torch.device("cpu"), sampled[lo:hi], np.diff(offsets.astype(np.int64)))
        has_n = np.array([(s.bases[int(o[i]):int(o[i + 1])] == ord("N")).any() for i in range(N_READS)])
        want = [i for i in range(lo) if sampled[i] and not has_n[i]]
        assert n_ctx == len(want)
        if want:
            cb, co = ctx.context
            assert np.array_equal(np.diff(co), [int(o[i + 1] - o[i]) for i in want])
            assert np.array_equal(cb, np.concatenate([s.bases[int(o[i]):int(o[i + 1])] for i in want]))
        else:
            assert ctx.context is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchanges_under_gloo(world):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, 11), nprocs=world, join=True)
