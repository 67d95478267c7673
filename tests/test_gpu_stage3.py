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


# ------------------------------------------------------------------------------------------------ DNA / edit-script stream
def _dna_round_trip(s, k, f, lo, hi, c, P, level, sparse_g=None, packs=None, is_hifi=False):
    n = s.n_reads
    with lib.Context(k, f, lo, hi, c, is_hifi=is_hifi) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        st = ctx.count_finalize()
        if sparse_g is None:
            sampled = np.ones(n, np.uint8)
        else:      # compression.cpp:443, :501-503
            mean_len = int(st["tot_kmers"] * f / max(1, st["n_reads"]) + k - 1)
            sampled = lib.sampler(max(1, int(sparse_g * st["n_unique_counted"] * f / max(1, mean_len))), 1.0, 0, n)
        ctx.graph_build(sampled)
        ctx.encode(P, packs)
        ctx.dna_encode(level, packs)
        stream, hdr = ctx.dna_stream()
        tuples = ctx.encode_size()
    has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(n)], np.uint8)
    is_ref = (sampled & (1 - has_n)).astype(np.uint8)
    bases, off = oracle_lib.dna_decode(stream, n, is_ref, s.n_bases)
    assert np.array_equal(off, s.offsets)
    assert np.array_equal(bases, s.bases)
    return len(stream), hdr, tuples


@pytest.mark.parametrize("level,prof", [(1, "ont"), (2, "ont"), (3, "clr")])
def test_dna_stream_round_trip(level, prof):
    """Device container -> independent C decoder -> the input reads, byte for byte (plain reads, reads with N, edit scripts,
    alternative reads, skips, every history width)."""
    s = synth.generate(800, 150000, 3000, seed=40 + level, profile=prof, n_frac=0.02)
    P = dict(P_BAL, min_part_len_alt=64 if level == 1 else 48, max_recurence=3 if level == 1 else 5)
    size, hdr, tuples = _dna_round_trip(s, 20, 9, 3, 100, 8, P, level, packs=[300, 1, 499])
    assert size < tuples


@pytest.mark.parametrize("level,prof,c", [(1, "ont", 5), (2, "ont", 8), (3, "clr", 10)])
def test_dna_stream_equals_cpu_twin(level, prof, c):
    """Device container == the oracle's twin encoder fed with the device's own (bit-exact) tuples, byte for byte — the same kind of
    pin the quality and header containers have.  (On the CPU the twin is checked on the reference's tuples through both decoders:
    tests/test_host_decode.py::test_reference_tuples_through_the_dna_container_on_cpu.)"""
    s = synth.generate(500, 120000, 2500, seed=70 + level, profile=prof, n_frac=0.02)
    P = dict(P_BAL, min_part_len_alt=64 if level == 1 else 48, max_recurence=3 if level == 1 else 5)
    packs = [200, 300]
    n = s.n_reads
    with lib.Context(20, 9, 3, 100, c) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        sampled = np.ones(n, np.uint8)
        ctx.graph_build(sampled)
        ctx.encode(P, packs)
        es_off, es = ctx.encoded(n)
        ctx.dna_encode(level, packs)
        stream, _ = ctx.dna_stream()
    has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(n)], np.uint8)
    is_ref = (sampled & (1 - has_n)).astype(np.uint8)
    es_list = [es[int(es_off[i]):int(es_off[i + 1])].tobytes() for i in range(n)]
    twin = oracle_lib.dna_encode(level, c, es_list, s.bases, s.offsets, is_ref, packs)
    assert len(twin) == len(stream) and np.array_equal(twin, stream)


def test_dna_stream_size_vs_reference():
    """12 500 synthetic ONT reads / 100 Mbases (BASELINE.md §2 recipe, seed 1) at the compress-ont default (k20 a16 f12 L4 H80 c5
    sparse g=1, level 1): the unmodified reference writes an 18 231 949-byte DNA stream.  The tuples are the
    reference's own (bit-exact stage 2), so the two coders see the same events."""
    s = synth.generate(12500, 5_000_000, 8000, seed=1, profile="ont")
    P = dict(anchor_len=16, k=20, modulo=12, hifi=0, min_part_len_alt=64, max_recurence=3, min_anchors=1,
             min_mmer_frac=0.5, min_mmer_force=0.9, max_matches_mult=10.0, es_cost_mult=1.0)
    size, hdr, tuples = _dna_round_trip(s, 20, 12, 4, 80, 5, P, 1, sparse_g=1.0)
    print(f"native DNA container {size} B (tables {hdr} B) vs reference 18231949 B: {size / 18231949:.4f}")
    assert size <= 1.005 * 18_231_949, (size, hdr)


# ------------------------------------------------------------------------------------------------ header stream
def _device_headers(hs, plus=None, packs=None):
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        ctx.hdr_encode(hs, plus, packs)
        return ctx.hdr_stream()


@pytest.mark.parametrize("packs", [None, "ragged"])
def test_header_stream_edges(packs):
    """Every branch of the event model: the device container equals the CPU twin byte for byte and decodes to the input."""
    from test_oracle_stage3 import hdr_edge_cases
    for case, hs in hdr_edge_cases().items():
        n = len(hs)
        ps = None
        if packs == "ragged" and n:
            ps = [min(n, 3), 0, max(0, n - 3)]
        plus = np.array([(i * 7) % 3 == 0 for i in range(n)], np.uint8)
        got, _ = _device_headers(hs, plus, ps)
        want = oracle_lib.hdr_encode(hs, plus, ps)
        assert np.array_equal(got, want), case
        dec, dplus = oracle_lib.hdr_decode(got, n, sum(map(len, hs)))
        assert dec == hs and np.array_equal(dplus, plus), case


def test_header_stream_fixtures_and_size():
    """Headers of the reference's own test inputs + 20 000 synthetic ONT headers: device == twin, decode == input, and the
    synthetic case is no larger than the unmodified reference's header stream for the same headers."""
    for name, (hs, ref_bytes) in golden_io.load_hdr_golden().items():
        got, tables = _device_headers(hs)
        assert np.array_equal(got, oracle_lib.hdr_encode(hs)), name
        dec, _ = oracle_lib.hdr_decode(got, len(hs), sum(map(len, hs)))
        assert dec == hs, name
        if name.startswith("synthetic"):
            print(f"native header container {len(got)} B (tables {tables} B) vs reference {ref_bytes} B")
            assert len(got) <= ref_bytes


def test_header_stream_refuses_nul():
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        with pytest.raises(lib.ClbError):
            ctx.hdr_encode([b"@a", b"@b\x00c"])


# ------------------------------------------------------------------------------------------------ lossless quality stream (-q org)
@pytest.mark.parametrize("name,source", [("ont", 0), ("hifi", 2), ("hifi", 1)])
def test_lossless_quality_stream_fixture(name, source):
    """Reference test inputs, level 1: device container == CPU twin byte for byte, decodes to the input qualities."""
    bases, quals, _, off = golden_io.load_qual_golden(name)
    n = len(off) - 1
    packs = [n // 2, n - n // 2]
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        ctx.append_reads(bases, off)
        ctx.count_finalize()
        ctx.qual_encode_original(source, 1, quals, off, packs)
        got = ctx.qual_stream()
    assert np.array_equal(got, oracle_lib.qorg_encode(source, 1, bases, quals, off, packs))
    assert np.array_equal(oracle_lib.qorg_decode(got, bases, off), quals)


@pytest.mark.parametrize("level,prof,source", [(2, "ont", 0), (3, "clr", 1), (2, "hifi", 2)])
def test_lossless_quality_stream_with_flags(level, prof, source):
    """Level > 1: match / anchor flags from the device's own tuples enter the context; N reads included."""
    s = synth.generate(500, 120000, 3000, seed=60 + level, profile=prof, n_frac=0.02)
    with lib.Context(20, 9, 3, 100, 8) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        ctx.graph_build(np.ones(s.n_reads, np.uint8))
        ctx.encode(P_BAL, [500])
        es_off, es = ctx.encoded(s.n_reads)
        ctx.qual_encode_original(source, level, s.quals, s.offsets, [500])
        got = ctx.qual_stream()
    assert np.array_equal(got, oracle_lib.qorg_encode(source, level, s.bases, s.quals, s.offsets, [500], es, es_off))
    assert np.array_equal(oracle_lib.qorg_decode(got, s.bases, s.offsets, es, es_off), s.quals)


def test_lossless_quality_stream_size_vs_reference():
    """2 000 synthetic HiFi reads / 29.9 Mbases (colord_b200.synth, seed 2): the unmodified reference's `compress-pbhifi -q org`
    writes a 19 286 491-byte quality stream for this input (oracle/_ref/colord, this container)."""
    s = synth.generate(2000, 3_000_000, 15000, seed=2, profile="hifi")
    with lib.Context(21, 40, 3, 80, 8, is_hifi=True) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        ctx.qual_encode_original(2, 1, s.quals, s.offsets)
        got = ctx.qual_stream()
    print(f"native lossless quality container {len(got)} B vs reference 19286491 B: {len(got) / 19286491:.4f}")
    assert len(got) <= 1.005 * 19_286_491
    assert np.array_equal(oracle_lib.qorg_decode(got, s.bases, s.offsets), s.quals)


def test_lossless_quality_refuses_bad_bytes():
    s = synth.generate(20, 20000, 1000, seed=9, profile="ont")
    q = s.quals.copy(); q[5] = 20
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        with pytest.raises(lib.ClbError):
            ctx.qual_encode_original(0, 1, q, s.offsets)
