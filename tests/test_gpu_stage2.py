"""Parity of the CUDA stage 2 (through the C-ABI) against the reference's dumps and the C oracle.  GPU only."""
import numpy as np
import pytest

import golden_io
import oracle_lib
from colord_b200 import lib

pytestmark = pytest.mark.gpu


def _ctx():
    return lib.Context(20, 12, 3, 80, 5)


def test_edit_scripts_golden():
    """E6/E7: every edit script equals what the reference's edlib wrappers + refactor_edit_script produced
    (tests/golden/edit_scripts.bin.gz: tiny inputs, flanks, sizes on edlib's Hirschberg path)."""
    cases = golden_io.load_edit_scripts()
    with _ctx() as ctx:
        got = ctx.edit_scripts([(k, r, e, rt, et) for k, r, e, rt, et, _ in cases])
    bad = [(i, c[0], len(c[1]), len(c[2])) for i, c in enumerate(cases) if got[i] != c[5]]
    assert not bad, (len(bad), bad[:10])


def _mutate(rng, s, rate):
    out = []
    for b in s:
        x = rng.random()
        if x < rate * 0.4:
            out.append((b + rng.integers(1, 4)) % 4)
        elif x < rate * 0.7:
            continue
        elif x < rate:
            out.append(b); out.append(rng.integers(0, 4))
        else:
            out.append(b)
    return np.array(out, np.uint8)


@pytest.mark.parametrize("seed", [1, 2])
def test_edit_scripts_random_vs_oracle(seed):
    """Random related / unrelated / homopolymer-rich pairs of every lane-group class against the C oracle."""
    rng = np.random.default_rng(seed)
    cases = []
    for n in [0, 1, 2, 3, 5, 14, 15, 16, 31, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256, 300, 511, 513, 700, 1100, 2500]:
        for kind in (0, 1, 2):
            for rate in (0.0, 0.1, 0.35):
                alpha = 2 if (n % 3 == 0) else 4
                ref = rng.integers(0, alpha, n).astype(np.uint8)
                enc = _mutate(rng, ref, rate) if n else rng.integers(0, 4, 3).astype(np.uint8)
                if kind != 2 and n > 4:        # flanks: unequal lengths, reference overhang
                    ref = np.concatenate([rng.integers(0, 4, n // 2).astype(np.uint8), ref]) if kind == 0 else np.concatenate([ref, rng.integers(0, 4, n // 2).astype(np.uint8)])
                cases.append((kind, ref, enc, int(rng.integers(0, 4)), 255))
    cases.append((2, np.zeros(0, np.uint8), np.zeros(0, np.uint8), 255, 255))
    with _ctx() as ctx:
        got = ctx.edit_scripts(cases)
    bad = []
    for i, (kind, ref, enc, rt, et) in enumerate(cases):
        want = oracle_lib.edit_script(ref, enc, kind, bytes([rt]), bytes([et]))
        if got[i] != want:
            bad.append((i, kind, len(ref), len(enc)))
    assert not bad, (len(bad), bad[:10])


# ------------------------------------------------------------------------------------------------ the whole encoder
from conftest import GOLDEN_CASES
from test_oracle_stage2 import golden_stage2_inputs

MMER_CASES = [c for c in GOLDEN_CASES if c != "hifi"]


def _run_stage2(g, keep_candidates=False, packs=True):
    p = g.params
    r = g.reads_in
    ctx = lib.Context(p["k"], p["modulo"], p["min_count"], p["max_count"], p["max_candidates"], is_hifi=bool(p.get("hifi", 0)))
    ctx.append_reads(r.bases, r.offsets)
    ctx.count_finalize()
    sampled = lib.sampler(p["sparse_range"], float(p["sparse_exponent"]), 0, r.n_reads) if p.get("sparse", 0) else np.ones(r.n_reads, np.uint8)
    ctx.graph_build(sampled)
    ctx.encode(p, np.array(g.packs, np.uint32) if packs else None, keep_candidates=keep_candidates)
    return ctx


@pytest.mark.parametrize("case", MMER_CASES)
def test_candidates_match_oracle(golden, case):
    """E1-E4: anchors of every read against every candidate = the oracle's (which is pinned on the reference's tuples)."""
    g = golden(case)
    cand, cand_n, _ = golden_stage2_inputs(g)
    want = oracle_lib.candidates(g.reads_in.bases, g.reads_in.offsets, g.is_ref, cand, cand_n, g.params)
    with _run_stage2(g, keep_candidates=True) as ctx:
        got = ctx.encode_candidates(g.reads_in.n_reads)
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, (len(bad), bad[:10], got[bad[0]][:1], want[bad[0]][:1])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_compact_es_bytes_golden(golden, case):
    """E1-E9: the CompactES bytes of every read equal what the unmodified reference's CEncoder produced."""
    g = golden(case)
    with _run_stage2(g) as ctx:
        off, es = ctx.encoded(g.reads_in.n_reads)
    got = [es[int(off[i]):int(off[i + 1])].tobytes() for i in range(g.reads_in.n_reads)]
    bad = [i for i in range(len(got)) if got[i] != g.es[i]]
    assert not bad, (len(bad), bad[:10])


@pytest.mark.parametrize("env", ["CLB_ALIGN_ONE_KERNEL", "CLB_ALIGN_THREAD_BACK", "CLB_SLAB_GB"])
@pytest.mark.parametrize("case", ["ont_bal", "clr_ratio"])
def test_compact_es_bytes_golden_kernel_variants(golden, case, env, monkeypatch):
    """The switchable forms of the alignment (one kernel per bin instead of forward + backward kernels; backward half with one thread
    per task, align_back.cuh) and the path without the device slab give the reference's CompactES bytes as well."""
    monkeypatch.setenv(env, "0" if env == "CLB_SLAB_GB" else "1")
    g = golden(case)
    with _run_stage2(g) as ctx:
        off, es = ctx.encoded(g.reads_in.n_reads)
    got = [es[int(off[i]):int(off[i + 1])].tobytes() for i in range(g.reads_in.n_reads)]
    bad = [i for i in range(len(got)) if got[i] != g.es[i]]
    assert not bad, (env, len(bad), bad[:10])


def _device_vs_oracle(s, k, f, lo, hi, c, P, pack_sizes=None, sampled=None):
    n = s.n_reads
    sampled = np.ones(n, np.uint8) if sampled is None else sampled
    with lib.Context(k, f, lo, hi, c) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        ctx.graph_build(sampled)
        cand, cn = ctx.graph_candidates()
        ctx.encode(P, pack_sizes)
        off, es = ctx.encoded(n)
    got = [es[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)]
    has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(n)], np.uint8)
    is_ref = (sampled & (1 - has_n)).astype(np.uint8)
    if pack_sizes is None:          # the reference's pack rule (in_reads.cpp:62-76)
        pack_sizes, acc, cnt = [], 0, 0
        for i in range(n):
            acc += int(s.offsets[i + 1] - s.offsets[i]) + 1; cnt += 1
            if acc >= (2 << 21):
                pack_sizes.append(cnt); acc = 0; cnt = 0
        if cnt:
            pack_sizes.append(cnt)
    want = oracle_lib.encode_reads(s.bases, s.offsets, is_ref, cand, cn, np.array(pack_sizes, np.uint32), P)
    bad = [i for i in range(n) if got[i] != want[i]]
    n_es = sum(1 for e in want if (e[0] >> 4) == 10)
    return bad, n_es


P_BAL = dict(anchor_len=16, k=20, modulo=9, hifi=0, min_part_len_alt=48, max_recurence=5, min_anchors=1,
             min_mmer_frac=0.5, min_mmer_force=0.9, max_matches_mult=10.0, es_cost_mult=1.0)


def test_encode_ont_many_packs_and_batches(monkeypatch):
    """1500 ONT-like reads at 24x: alternative-read recursion, reads with N, several packs, several device batches."""
    from colord_b200 import synth
    monkeypatch.setenv("CLB_BATCH_MBASES", "2")
    monkeypatch.setenv("CLB_ALIGN_SCRATCH_MB", "64")
    s = synth.generate(1500, 250000, 4000, seed=5, profile="ont", n_frac=0.02)
    packs = [100, 1, 399, 500, 250, 250]
    bad, n_es = _device_vs_oracle(s, 20, 9, 3, 100, 8, P_BAL, packs)
    assert n_es > 1300
    assert not bad, (len(bad), bad[:10])


def test_encode_long_reads():
    """60 kb reads: the m-mer table and Bloom filter leave shared memory, flanks reach edlib's Hirschberg sizes."""
    from colord_b200 import synth
    s = synth.generate(100, 500000, 60000, seed=6, profile="ont")
    bad, n_es = _device_vs_oracle(s, 20, 9, 2, 100, 8, P_BAL)
    assert n_es > 90
    assert not bad, (len(bad), bad[:10])


def test_encode_accurate_reads_big_segments():
    """0.2 % error: thousands of m-mer pairs per candidate (global-memory sort), long anchors, few edit operations."""
    from colord_b200 import synth
    s = synth.generate(300, 100000, 6000, seed=7, profile="hifi")
    bad, n_es = _device_vs_oracle(s, 20, 9, 2, 100, 8, P_BAL)
    assert n_es > 280
    assert not bad, (len(bad), bad[:10])


def test_encode_repeats():
    """Tandem repeats, a microsatellite and a homopolymer: duplicate m-mers on both sides, the match-count cap, overlap fixes."""
    from colord_b200 import synth
    rng = np.random.default_rng(11)
    unit = rng.integers(0, 4, 1500, dtype=np.uint8)
    parts = [rng.integers(0, 4, 20000, dtype=np.uint8), unit, unit, unit, rng.integers(0, 4, 8000, dtype=np.uint8),
             np.tile(np.array([0, 1, 2, 3, 3], np.uint8), 300), rng.integers(0, 4, 6000, dtype=np.uint8), np.zeros(400, np.uint8),
             rng.integers(0, 4, 15000, dtype=np.uint8), unit, rng.integers(0, 4, 5000, dtype=np.uint8)]
    genome = np.concatenate(parts)
    seqs, offs = [], [0]
    for i in range(500):
        ln = int(np.clip(rng.gamma(2.0, 2500), 300, 20000))
        st = int(rng.integers(0, len(genome) - 300))
        frag = genome[st:st + ln].copy()
        if rng.random() < 0.5:
            frag = 3 - frag[::-1]
        r = rng.random(len(frag))
        sub = r < 0.02
        frag[sub] = (frag[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
        frag = frag[~((r >= 0.02) & (r < 0.04))]
        seqs.append(np.frombuffer(b"ACGT", np.uint8)[frag]); offs.append(offs[-1] + len(frag))
    s = synth.SynthReads(np.concatenate(seqs), None, np.array(offs, np.uint64), None)
    P = dict(P_BAL, anchor_len=14)
    bad, n_es = _device_vs_oracle(s, 18, 5, 2, 200, 10, P)
    assert n_es > 400
    assert not bad, (len(bad), bad[:10])


def test_encode_hifi_kmer_anchors_vs_oracle():
    """HiFi path (encoder.cpp:870-1012, :1113-1147, :1194-1253): anchors from the graph's shared k-mers, m-mer fallback,
    against the C oracle fed with the device's own candidates and shared k-mers."""
    from colord_b200 import synth
    s = synth.generate(400, 150000, 6000, seed=31, profile="hifi", n_frac=0.01)
    n = s.n_reads
    k, f, lo, hi, c = 21, 20, 2, 100, 8
    P = dict(P_BAL, anchor_len=18, k=k, modulo=f, hifi=1)
    sampled = np.ones(n, np.uint8)
    with lib.Context(k, f, lo, hi, c, is_hifi=True) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        ctx.graph_build(sampled)
        cand, cn = ctx.graph_candidates()
        common = ctx.graph_common()
        ctx.encode(P, [150, 250])
        off, es = ctx.encoded(n)
    got = [es[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)]
    has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(n)], np.uint8)
    is_ref = (sampled & (1 - has_n)).astype(np.uint8)
    want = oracle_lib.encode_reads(s.bases, s.offsets, is_ref, cand, cn, np.array([150, 250], np.uint32), P, common)
    bad = [i for i in range(n) if got[i] != want[i]]
    assert sum(1 for e in want if (e[0] >> 4) == 10) > 350
    assert not bad, (len(bad), bad[:10])
