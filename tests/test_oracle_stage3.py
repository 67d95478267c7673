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
