"""Pin the quality-stream oracle (oracle/stage3_qual.c) against the reference's own lossy-quality fixtures (test/*.quan,
copied into tests/golden/qual_*.bin.gz by tests/golden/make_qual_golden.py), and check the native container's round trip.
CPU only.  The native container has no reference bitstream to compare with: its parity is (a) identical reconstructed
qualities, (b) stream size against the reference's own qual stream (tests/golden/sizes.json, GPU test at full size)."""
import numpy as np
import pytest

import golden_io
import oracle_lib

CASES = {"ont": (4, [7, 14, 26]), "hifi": (5, [7, 14, 26, 93])}


@pytest.mark.parametrize("name", list(CASES))
def test_lossy_transform_equals_reference_quan(name):
    n_bins, thr = CASES[name]
    bases, quals, quan, off = golden_io.load_qual_golden(name)
    got = oracle_lib.qual_lossy(oracle_lib.qual_params(n_bins, thr, 1), bases, quals, off)
    assert np.array_equal(got, quan)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("packs", [1, 3])
def test_native_container_round_trip(name, packs):
    n_bins, thr = CASES[name]
    bases, quals, quan, off = golden_io.load_qual_golden(name)
    n = len(off) - 1
    cuts = np.linspace(0, n, packs + 1).astype(int)
    P = oracle_lib.qual_params(n_bins, thr, 1)
    stream = oracle_lib.qual_encode(P, bases, quals, off, np.diff(cuts))
    assert np.array_equal(oracle_lib.qual_decode(stream, bases, off), quan)


# ------------------------------------------------------------------------------------------------ header stream
def hdr_edge_cases():
    """Header lists exercising every branch of the event model (oracle/stage3_hdr.c)."""
    long_tok = b"@" + b"x" * 300 + b" " + b"y" * 40
    return {
        "empty_list": [],
        "single": [b"@only one"],
        "empty_strings": [b"", b"", b"@a", b"", b"@a"],
        "shape_changes": [b"@a.1 x=1", b"@a.2 x=22", b"@b-3", b"@b-4", b"@a.5 x=333", b"@a.5 x=333", b"@a.6/x=333"],
        "long_tokens": [long_tok, long_tok[:-1] + b"z", long_tok + b"q", b"@" + b"x" * 299 + b" " + b"y" * 41],
        "many_tokens": [b":".join(b"%d" % (i + k) for k in range(90)) for i in range(40)],
        "separators_only": [b"///", b"///", b"//", b"/:/", b"/:/"],
        "high_bytes": [bytes([200, 201, 32, 250]), bytes([200, 202, 32, 251]), bytes([200, 202, 32, 251])],
        "counter": [b"@r%d" % i for i in range(5000)],
    }


@pytest.mark.parametrize("case", list(hdr_edge_cases()))
@pytest.mark.parametrize("packs", [None, "ragged"])
def test_header_container_round_trip_edges(case, packs):
    hs = hdr_edge_cases()[case]
    n = len(hs)
    ps = None
    if packs == "ragged":
        ps = [p for p in [1, 0, 2, 64, 65] if p <= n]
        ps = ps[:next((i for i in range(len(ps) + 1) if sum(ps[:i + 1]) > n), len(ps))]
        ps.append(n - sum(ps))
    plus = np.array([(i * 7) % 3 == 0 for i in range(n)], np.uint8)
    stream = oracle_lib.hdr_encode(hs, plus, ps)
    dec, dplus = oracle_lib.hdr_decode(stream, n, sum(map(len, hs)))
    assert dec == hs
    assert np.array_equal(dplus, plus)


def test_header_container_fixtures_and_size():
    """Headers of the reference's own test files round-trip; on 20 000 synthetic ONT headers the container is smaller than the
    header stream the unmodified reference wrote for the same headers (tests/golden/make_hdr_golden.py)."""
    for name, (hs, ref_bytes) in golden_io.load_hdr_golden().items():
        stream = oracle_lib.hdr_encode(hs)
        dec, _ = oracle_lib.hdr_decode(stream, len(hs), sum(map(len, hs)))
        assert dec == hs, name
        if name.startswith("synthetic"):
            assert len(stream) <= ref_bytes, (len(stream), ref_bytes)


# ------------------------------------------------------------------------------------------------ lossless quality stream (-q org)
# quality-stream bytes the unmodified reference wrote with `-q org` for its own test files (ONT 284 680, HiFi 397 555)
REF_QORG_BYTES = {"ont": (0, 284_680), "hifi": (2, 397_555)}


@pytest.mark.parametrize("name", list(REF_QORG_BYTES))
@pytest.mark.parametrize("level", [1, 3])
def test_lossless_quality_container_round_trip(name, level):
    """The reference's own test inputs: the container decodes to the input qualities (every source's quantiser, level 1 and the
    wider level-3 context) and, at level 1, is no larger than the reference's quality stream for the same file."""
    source, ref_bytes = REF_QORG_BYTES[name]
    bases, quals, _, off = golden_io.load_qual_golden(name)
    n = len(off) - 1
    es = es_off = None
    if level > 1:      # flags need tuples: plain-read tuples (no match / anchor flags) are enough for the round trip
        es_off = np.zeros(n + 1, np.uint64); es_off[1:] = np.cumsum(np.ones(n, np.uint64))
        es = np.full(n, 9 << 4, np.uint8)
    for src in ({source, 1} if level == 1 else {source}):
        stream = oracle_lib.qorg_encode(src, level, bases, quals, off, [n // 3, 1, n - n // 3 - 1], es, es_off)
        assert np.array_equal(oracle_lib.qorg_decode(stream, bases, off, es, es_off), quals)
        if level == 1 and src == source:
            assert len(stream) <= ref_bytes, (len(stream), ref_bytes)
