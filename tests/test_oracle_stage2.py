"""Pin the plain-C restatement of stage 2 (oracle/stage2.c) against the CompactES bytes the unmodified reference's
CEncoder produced (tests/golden/<case>/es.bin.gz, written by oracle/_ref/ref_stage_dump).  CPU only."""
import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN_CASES


def golden_stage2_inputs(g):
    p = g.params
    n = len(g.reads)
    mc = p["max_candidates"]
    cand = np.zeros((n, mc), np.uint32)
    cand_n = np.zeros(n, np.uint32)
    for i, r in enumerate(g.reads):
        cand_n[i] = len(r["cands"])
        cand[i, :len(r["cands"])] = r["cands"]
    common = None
    if p["hifi"]:
        coff = np.zeros(n * mc, np.uint64)
        cn = np.zeros(n * mc, np.uint32)
        parts, tot = [], 0
        for i, r in enumerate(g.reads):
            for j, km in enumerate(r["common"]):
                coff[i * mc + j] = tot
                cn[i * mc + j] = len(km)
                parts.append(km)
                tot += len(km)
        common = (coff, cn, np.concatenate(parts).astype(np.uint64) if parts else np.zeros(0, np.uint64))
    return cand, cand_n, common


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_compact_es_bytes(golden, case):
    g = golden(case)
    cand, cand_n, common = golden_stage2_inputs(g)
    assert g.es_packs == g.packs
    got = oracle_lib.encode_reads(g.reads_in.bases, g.reads_in.offsets, g.is_ref, cand, cand_n, np.array(g.packs, np.uint32), g.params, common)
    bad = [i for i in range(len(got)) if got[i] != g.es[i]]
    assert not bad, (len(bad), bad[:10])


def test_edit_scripts_match_reference():
    """E6/E7 in isolation, incl. tiny inputs, flanks and inputs large enough for edlib's Hirschberg path."""
    import golden_io
    cases = golden_io.load_edit_scripts()
    assert len(cases) > 300
    bad = []
    for i, (kind, ref, enc, rt, et, want) in enumerate(cases):
        got = oracle_lib.edit_script(ref, enc, kind, bytes([rt]), bytes([et]))
        if got != want:
            bad.append((i, kind, len(ref), len(enc)))
    assert not bad, bad[:10]
