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


@pytest.mark.parametrize("case", MMER_CASES)
def test_compact_es_bytes_golden(golden, case):
    """E1-E9: the CompactES bytes of every read equal what the unmodified reference's CEncoder produced."""
    g = golden(case)
    with _run_stage2(g) as ctx:
        off, es = ctx.encoded(g.reads_in.n_reads)
    got = [es[int(off[i]):int(off[i + 1])].tobytes() for i in range(g.reads_in.n_reads)]
    bad = [i for i in range(len(got)) if got[i] != g.es[i]]
    assert not bad, (len(bad), bad[:10])
