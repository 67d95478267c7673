"""Pin the plain-C restatement (oracle/stage1.c) against dumps of the unmodified reference.

The dumps under tests/golden/ were written by oracle/_ref/ref_stage_dump (reference classes with taps,
see oracle/ref_stage_dump.cpp and tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN_CASES


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_count_and_filter(golden, case):
    g = golden(case)
    p = g.params
    km, ct, st = oracle_lib.count_kmers(g.reads_in.bases, g.reads_in.offsets, p["k"], p["modulo"], p["min_count"], p["max_count"])
    order = np.argsort(g.kmers, kind="stable")
    assert np.array_equal(km, g.kmers[order])
    assert np.array_equal(ct, g.counts[order])
    assert st["n_reads"] == p["n_reads"]
    assert st["tot_kmers"] == p["tot_kmers"]
    assert st["n_unique_counted"] == p["n_unique_counted"]
    assert st["total_count_filtered"] == p["total_count_filtered"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_accepted_kmers(golden, case):
    g = golden(case)
    p = g.params
    off, acc = oracle_lib.accepted_kmers(g.reads_in.bases, g.reads_in.offsets, p["k"], p["modulo"], np.sort(g.kmers))
    assert len(off) - 1 == len(g.reads)
    for i, r in enumerate(g.reads):
        assert r["id"] == i
        assert np.array_equal(acc[int(off[i]):int(off[i + 1])], r["acc"]), i


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_sampler(golden, case):
    g = golden(case)
    p = g.params
    n = len(g.reads)
    if p["sparse"]:
        dec = oracle_lib.sampler(p["sparse_range"], float(p["sparse_exponent"]), 0, n)
    else:
        dec = np.ones(n, np.uint8)
    assert np.array_equal(dec & (1 - g.has_n), g.is_ref)
    assert int((dec if p["sparse"] else np.ones(n, np.uint8)).sum()) == p["tot_ref_reads"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_sim_graph(golden, case):
    g = golden(case)
    p = g.params
    n = len(g.reads)
    acc_off = np.zeros(n + 1, np.uint64)
    acc_off[1:] = np.cumsum([len(r["acc"]) for r in g.reads])
    acc = np.concatenate([r["acc"] for r in g.reads]).astype(np.uint64)
    sampled = oracle_lib.sampler(p["sparse_range"], float(p["sparse_exponent"]), 0, n) if p["sparse"] else np.ones(n, np.uint8)
    cand, cand_n, common = oracle_lib.sim_graph(acc_off, acc, g.has_n, sampled, p["max_candidates"], p["max_count"], hifi=bool(p["hifi"]))
    for i, r in enumerate(g.reads):
        assert np.array_equal(cand[i, :cand_n[i]], r["cands"]), i
        if p["hifi"]:
            coff, cn, cm = common
            assert len(r["common"]) == cand_n[i]
            for j in range(cand_n[i]):
                slot = i * p["max_candidates"] + j
                assert np.array_equal(cm[int(coff[slot]):int(coff[slot]) + int(cn[slot])], r["common"][j]), (i, j)
