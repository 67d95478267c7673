"""Parity of the CUDA stage 1 (through the C-ABI) against the reference dumps and the C oracle.  GPU only."""
import numpy as np
import pytest

import oracle_lib
from colord_b200 import lib, synth
from conftest import GOLDEN_CASES

pytestmark = pytest.mark.gpu


def _ctx(p, **kw):
    return lib.Context(p["k"], p["modulo"], p["min_count"], p["max_count"], p["max_candidates"], is_hifi=bool(p.get("hifi", 0)), **kw)


def _sampled(p, n):
    if p.get("sparse", 0):
        return lib.sampler(p["sparse_range"], float(p["sparse_exponent"]), 0, n)
    return np.ones(n, np.uint8)


def _check_filter(ctx, stats, want_kmers, want_counts, want_stats):
    km, ct = ctx.filter_list()
    o = np.argsort(km, kind="stable")
    w = np.argsort(want_kmers, kind="stable")
    assert np.array_equal(km[o], want_kmers[w])
    assert np.array_equal(ct[o], want_counts[w])
    for k, v in want_stats.items():
        assert stats[k] == v, k


def _check_graph(ctx, want_acc_off, want_acc, want_cand, want_cand_n, want_common=None):
    off, acc = ctx.graph_accepted()
    assert np.array_equal(off, want_acc_off)
    assert np.array_equal(acc, want_acc)
    cand, cn = ctx.graph_candidates()
    assert np.array_equal(cn, want_cand_n)
    for i in range(len(cn)):
        assert np.array_equal(cand[i, :cn[i]], want_cand[i, :cn[i]]), i
    if want_common is not None:
        coff, ccn, ckm = ctx.graph_common()
        woff, wcn, wkm = want_common
        mc = cand.shape[1]
        for i in range(len(cn)):
            for j in range(cn[i]):
                s = i * mc + j
                assert ccn[s] == wcn[s], (i, j)
                assert np.array_equal(ckm[int(coff[s]):int(coff[s]) + int(ccn[s])], wkm[int(woff[s]):int(woff[s]) + int(wcn[s])]), (i, j)


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("n_appends", [1, 3])
def test_golden_stage1(golden, case, n_appends):
    """Every stage-1 product equals what the unmodified reference produced (tests/golden)."""
    g = golden(case)
    p = g.params
    r = g.reads_in
    with _ctx(p) as ctx:
        n = r.n_reads
        cuts = np.linspace(0, n, n_appends + 1).astype(int)
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.append_reads(r.bases, r.offsets[a:b + 1])
        st = ctx.count_finalize()
        _check_filter(ctx, st, g.kmers, g.counts, dict(n_reads=p["n_reads"], tot_kmers=p["tot_kmers"], n_unique_counted=p["n_unique_counted"],
                                                       total_count_filtered=p["total_count_filtered"]))
        ctx.graph_build(_sampled(p, n))
        acc_off = np.zeros(n + 1, np.uint64)
        acc_off[1:] = np.cumsum([len(x["acc"]) for x in g.reads])
        acc = np.concatenate([x["acc"] for x in g.reads]).astype(np.uint64)
        mc = p["max_candidates"]
        want_cand = np.zeros((n, mc), np.uint32)
        want_cn = np.zeros(n, np.uint32)
        for i, x in enumerate(g.reads):
            want_cn[i] = len(x["cands"])
            want_cand[i, :len(x["cands"])] = x["cands"]
        _check_graph(ctx, acc_off, acc, want_cand, want_cn)
        if p["hifi"]:
            coff, ccn, ckm = ctx.graph_common()
            for i, x in enumerate(g.reads):
                for j in range(len(x["cands"])):
                    s = i * mc + j
                    assert np.array_equal(ckm[int(coff[s]):int(coff[s]) + int(ccn[s])], x["common"][j]), (i, j)
        # reference-read store layout (reference_reads.h:35-72)
        for i in (0, 1, n - 1):
            if not g.has_n[i]:
                b = r.bases[int(r.offsets[i]):int(r.offsets[i + 1])]
                want = np.zeros(len(b) // 4 + 2, np.uint8)
                m = oracle_lib.lib().orc_pack_ref_read(np.ascontiguousarray(b), len(b), want)
                assert np.array_equal(ctx.packed_read(i), want[:m])
        assert ctx.kernel_launches > 0


def _oracle_stage1(s, p, sampled):
    km, ct, st = oracle_lib.count_kmers(s.bases, s.offsets, p["k"], p["modulo"], p["min_count"], p["max_count"])
    off, acc = oracle_lib.accepted_kmers(s.bases, s.offsets, p["k"], p["modulo"], km)
    has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(s.n_reads)], np.uint8)
    cand, cn, common = oracle_lib.sim_graph(off, acc, has_n, sampled, p["max_candidates"], p["max_count"], hifi=bool(p.get("hifi", 0)))
    return km, ct, st, off, acc, cand, cn, common


def _run_vs_oracle(s, p, sampled, **kw):
    km, ct, st, off, acc, cand, cn, common = _oracle_stage1(s, p, sampled)
    with _ctx(p, **kw) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        got = ctx.count_finalize()
        _check_filter(ctx, got, km, ct, st)
        ctx.graph_build(sampled)
        _check_graph(ctx, off, acc, cand, cn, common)
    return st


PRESETS = {
    "ont_mem": dict(k=20, modulo=12, min_count=4, max_count=80, max_candidates=5, sparse=1),
    "ont_ratio": dict(k=21, modulo=8, min_count=2, max_count=120, max_candidates=10, sparse=0),
    "hifi_bal": dict(k=23, modulo=30, min_count=3, max_count=120, max_candidates=10, sparse=1, hifi=1),
    "k32": dict(k=32, modulo=5, min_count=2, max_count=50, max_candidates=8, sparse=0),
    "k15_f1": dict(k=15, modulo=1, min_count=2, max_count=30, max_candidates=4, sparse=0),
}


@pytest.mark.parametrize("preset", list(PRESETS))
def test_random_vs_oracle(preset):
    """Seeded synthetic reads (with N reads and short reads) at oracle-friendly size: everything bit-exact."""
    p = PRESETS[preset]
    prof = "hifi" if p.get("hifi") else "ont"
    s = synth.generate(600, 50000, 2500, seed=101, profile=prof, n_frac=0.05, min_len=10)
    sampled = lib.sampler(9, 1.0, 0, s.n_reads) if p.get("sparse") else np.ones(s.n_reads, np.uint8)
    st = _run_vs_oracle(s, p, sampled)
    assert st["n_unique_counted"] > 100


def test_table_growth_and_hint():
    """The count table grows on demand (no hint) and gives the same result with an exact hint."""
    p = PRESETS["ont_ratio"]
    s = synth.generate(1500, 200000, 3000, seed=7, profile="ont")
    sampled = np.ones(s.n_reads, np.uint8)
    _run_vs_oracle(s, p, sampled)
    _run_vs_oracle(s, p, sampled, expected_bases=s.n_bases)


def test_edge_cases():
    p = PRESETS["ont_ratio"]
    # empty input
    with _ctx(p) as ctx:
        st = ctx.count_finalize()
        assert st["n_unique_counted"] == 0 and st["tot_kmers"] == 0
        ctx.graph_build(None)
        cand, cn = ctx.graph_candidates()
        assert len(cn) == 0
    # empty reads, reads shorter than k, a read that is all N, reads of exactly k
    rng = np.random.default_rng(5)
    core = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 3000)]
    reads = [core[:0], core[:5], core[:p["k"]], core[:p["k"]], np.full(50, ord("N"), np.uint8), core[100:1500], core[100:1500], core[:0],
             core[200:2900], core[1000:3000], core[:p["k"] - 1]]
    bases = np.concatenate(reads)
    offsets = np.zeros(len(reads) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(x) for x in reads])
    s = synth.SynthReads(bases, bases.copy(), offsets, [b"x"] * len(reads))
    _run_vs_oracle(s, p, np.ones(len(reads), np.uint8))


def test_bad_symbol_is_an_error():
    p = PRESETS["ont_mem"]
    bases = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTacgtACGT", np.uint8)
    with _ctx(p) as ctx:
        ctx.append_reads(bases, np.array([0, len(bases)], np.uint64))
        with pytest.raises(lib.ClbError) as e:
            ctx.count_finalize()
        assert e.value.status == 4


def test_long_reads_use_larger_scratch_classes():
    """Reads beyond 65 536 and 262 144 bases go through the larger shared / global scratch classes of k_accept."""
    p = dict(k=20, modulo=3, min_count=2, max_count=80, max_candidates=5, sparse=0)
    rng = np.random.default_rng(9)
    genome = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 400000)]
    spans = [(0, 300000), (1000, 301000), (50000, 150000), (50500, 150100), (100, 4000), (0, 400000), (390000, 400000)]
    reads = [genome[a:b] for a, b in spans]
    bases = np.concatenate(reads)
    offsets = np.zeros(len(reads) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(x) for x in reads])
    s = synth.SynthReads(bases, bases.copy(), offsets, [b"x"] * len(reads))
    _run_vs_oracle(s, p, np.ones(len(reads), np.uint8))


def test_many_neighbours_and_capped_lists():
    """Every read shares a repeat: lists exceed max_count (sorted + truncated to the first ids) and a read has
    more distinct neighbours than the shared vote table holds (global scratch class of k_vote)."""
    p = dict(k=16, modulo=2, min_count=2, max_count=40, max_candidates=6, sparse=0)
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    repeat = acgt[rng.integers(0, 4, 600)]
    reads = []
    for i in range(3600):
        piece = repeat[(i * 7) % 300:][:120 + (i % 5)] if i < 3500 else repeat
        reads.append(np.concatenate([acgt[rng.integers(0, 4, 30)], piece, acgt[rng.integers(0, 4, 30)]]))
    # low max_count makes almost every repeat k-mer list overflow; the last 100 reads see > 3072 neighbours
    p2 = dict(p, max_count=4000)
    bases = np.concatenate(reads)
    offsets = np.zeros(len(reads) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(x) for x in reads])
    s = synth.SynthReads(bases, bases.copy(), offsets, [b"x"] * len(reads))
    _run_vs_oracle(s, p, np.ones(len(reads), np.uint8))
    _run_vs_oracle(s, p2, np.ones(len(reads), np.uint8))


def test_partitioned_exchange_equals_single_context():
    """The multi-GPU exchange (export by owner partition -> merge -> finalize share -> import union) on two
    contexts of one GPU reproduces the single-context filter and graph."""
    p = PRESETS["ont_ratio"]
    s = synth.generate(500, 40000, 2500, seed=21, profile="ont", n_frac=0.02)
    km, ct, st, off, acc, cand, cn, common = _oracle_stage1(s, p, np.ones(s.n_reads, np.uint8))
    half = s.n_reads // 2
    shards = [(0, half), (half, s.n_reads)]
    ctxs = [_ctx(p) for _ in shards]
    try:
        for c, (a, b) in zip(ctxs, shards):
            c.append_reads(s.bases, s.offsets[a:b + 1])
        exported = [[c.counts_export(part, 2) for part in range(2)] for c in ctxs]
        # the one-pass form of the same export (clb_counts_sizes / clb_counts_export_all: what the exchanges call) gives the same sets
        import torch
        for c, per_part in zip(ctxs, exported):
            for n_parts in (2, 5):
                sizes = c.counts_sizes(n_parts)
                first = [sum(sizes[:q]) for q in range(n_parts)]
                tot_n = sum(sizes)
                dk = torch.zeros(max(1, tot_n), dtype=torch.int64, device="cuda"); dc = torch.zeros(max(1, tot_n), dtype=torch.int32, device="cuda")
                c.counts_export_all_device(n_parts, first, dk.data_ptr(), dc.data_ptr(), max(1, tot_n))
                hk, hc = dk.cpu().numpy().view(np.uint64), dc.cpu().numpy().view(np.uint32)
                for q in range(n_parts):
                    k1, c1 = c.counts_export(q, n_parts)
                    assert sizes[q] == len(k1)
                    a, b = np.argsort(hk[first[q]:first[q] + sizes[q]]), np.argsort(k1)
                    assert np.array_equal(hk[first[q]:first[q] + sizes[q]][a], k1[b]) and np.array_equal(hc[first[q]:first[q] + sizes[q]][a], c1[b])
        for r, c in enumerate(ctxs):
            c.counts_reset()
            for src in range(2):
                k2, c2 = exported[src][r]
                c.counts_merge(k2, c2, n_reads_remote=(shards[src][1] - shards[src][0]) if src != r else 0)
        stats = [c.count_finalize() for c in ctxs]
        lists = [c.filter_list() for c in ctxs]
        allk = np.concatenate([l[0] for l in lists])
        allc = np.concatenate([l[1] for l in lists])
        tot = {k: sum(x[k] for x in stats) for k in ("tot_kmers", "n_unique", "n_unique_counted", "total_count_filtered")}
        tot["n_reads"] = s.n_reads
        for k in ("tot_kmers", "n_unique_counted", "total_count_filtered"):
            assert tot[k] == st[k], k
        o = np.argsort(allk)
        assert np.array_equal(allk[o], km) and np.array_equal(allc[o], ct)
        # rank 0 imports the union and builds the graph for its shard only
        c0 = ctxs[0]
        c0.filter_import(allk, allc, tot)
        c0.graph_build(np.ones(half, np.uint8))
        off0, acc0 = c0.graph_accepted()
        assert np.array_equal(off0, off[:half + 1])
        assert np.array_equal(acc0, acc[:int(off[half])])
        cand0, cn0 = c0.graph_candidates()
        assert np.array_equal(cn0, cn[:half])
    finally:
        for c in ctxs:
            c.close()
