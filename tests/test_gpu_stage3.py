"""Parity of the CUDA quality stream (through the C-ABI) against the oracle twin and the reference's fixtures.  GPU only."""
import numpy as np
import pytest

import golden_io
import oracle_lib
from colord_b200 import lib, synth

pytestmark = pytest.mark.gpu

CASES = {"ont": (4, [7, 14, 26]), "hifi": (5, [7, 14, 26, 93])}


def _device_stream(bases, quals, off, n_bins, thr, level, packs=None, encode_params=None):
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        ctx.append_reads(bases, off)
        ctx.count_finalize()
        es = es_off = None
        if level > 1:
            ctx.graph_build(np.ones(len(off) - 1, np.uint8))
            ctx.encode(encode_params, packs)
            es_off, es = ctx.encoded(len(off) - 1)
        ctx.qual_encode(n_bins, thr, level, quals, off, packs)
        return ctx.qual_stream(), es, es_off


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("packs", [None, [7, 1, 20, 12]])
def test_quality_stream_fixture(name, packs):
    """Reference fixtures: the device container equals the oracle twin byte for byte and decodes to the reference's .quan."""
    n_bins, thr = CASES[name]
    bases, quals, quan, off = golden_io.load_qual_golden(name)
    if packs is not None:
        n = sum(packs)
        bases, quals, quan, off = bases[:int(off[n])], quals[:int(off[n])], quan[:int(off[n])], off[:n + 1]
    got, _, _ = _device_stream(bases, quals, off, n_bins, thr, 1, packs)
    want = oracle_lib.qual_encode(oracle_lib.qual_params(n_bins, thr, 1), bases, quals, off, packs if packs is not None else [len(off) - 1])
    assert np.array_equal(got, want)
    assert np.array_equal(oracle_lib.qual_decode(got, bases, off), quan)


P_BAL = dict(anchor_len=16, k=20, modulo=9, hifi=0, min_part_len_alt=48, max_recurence=5, min_anchors=1,
             min_mmer_frac=0.5, min_mmer_force=0.9, max_matches_mult=10.0, es_cost_mult=1.0)


def test_quality_stream_level2_flags():
    """Level > 1: the match / anchor flags of the context come from the device's own tuples (quality_coder_impl.cpp:25-76)."""
    s = synth.generate(600, 120000, 3000, seed=21, profile="ont", n_frac=0.02)
    got, es, es_off = _device_stream(s.bases, s.quals, s.offsets, 4, [7, 14, 26], 2, [600], P_BAL)
    want = oracle_lib.qual_encode(oracle_lib.qual_params(4, [7, 14, 26], 2), s.bases, s.quals, s.offsets, [600], es, es_off)
    assert np.array_equal(got, want)
    lossy = oracle_lib.qual_lossy(oracle_lib.qual_params(4, [7, 14, 26], 2), s.bases, s.quals, s.offsets)
    assert np.array_equal(oracle_lib.qual_decode(got, s.bases, s.offsets, es, es_off), lossy)


def test_quality_stream_size_vs_reference():
    """Size parity at the north star's 0.5 %: 12 500 synthetic ONT reads / 100 Mbases (BASELINE.md §2 recipe, seed 1), for
    which the unmodified reference's `compress-ont` default writes an 18 452 133-byte qual stream."""
    s = synth.generate(12500, 5_000_000, 8000, seed=1, profile="ont")
    got, _, _ = _device_stream(s.bases, s.quals, s.offsets, 4, [7, 14, 26], 1)
    assert len(got) <= 1.005 * 18_452_133, len(got)
    lossy = oracle_lib.qual_lossy(oracle_lib.qual_params(4, [7, 14, 26], 1), s.bases, s.quals, s.offsets)
    assert np.array_equal(oracle_lib.qual_decode(got, s.bases, s.offsets), lossy)
