"""Compat streams (colord_b200/csrc/stage3_exact.cu) against the STOCK reference binary: the device's "dna", "qual" and "header" parts
must be the bytes the unmodified `colord compress-*` wrote into its archive for the same input (tests/golden/streams.json: size and
SHA-1 of every part; generator tests/golden/make_stream_golden.py), and equal to the CPU restatement oracle/stage3_exact.c on inputs
the goldens do not cover.  GPU only; everything goes through the C-ABI."""
import hashlib

import numpy as np
import pytest

import golden_io
import oracle_lib
from colord_b200 import lib, synth
from test_gpu_stage2 import _run_stage2
from test_oracle_exact import ONT_MEM, QUAL, STREAMS, synth_case

pytestmark = pytest.mark.gpu


def check(parts, want):
    assert len(parts) == len(want), (len(parts), len(want))
    for i, (p, (_, size, sha)) in enumerate(zip(parts, want)):
        assert len(p) == size, (i, len(p), size)
        assert hashlib.sha1(p).hexdigest() == sha, i


@pytest.mark.parametrize("name", ["ont_mem", "ont_bal", "clr_ratio", "hifi"])
def test_dna_parts_equal_the_reference(golden, name):
    """Stages 1 + 2 + the DNA coder on the device -> the stock binary's `dna` parts (levels 1, 2, 3; sparse and all-reference modes)."""
    g = golden(name)
    with _run_stage2(g) as ctx:
        parts = ctx.xdna_encode(g.params["level"], g.es_packs)
    check(parts, STREAMS[name]["streams"]["dna"])


@pytest.mark.parametrize("name", list(QUAL))
def test_quality_parts_equal_the_reference(golden, name):
    """Every -q mode of the reference (org for the three data sources, 2/4/5-avg, 2/4/5-fix, avg, none; level > 1 with the flags
    from the device's own tuples)."""
    mode, thr, source, level = QUAL[name]
    if level > 1:
        g = golden({"q_hifi_org": "hifi", "q_org_bal": "ont_bal"}.get(name, name))
        s = g.reads_in
        ctx = _run_stage2(g)
    else:
        s = synth_case(name)
        ctx = lib.Context(20, 12, 3, 80, 5)
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
    with ctx:
        parts = ctx.xqual_encode(mode, source, level, thr, s.quals, s.offsets, [s.n_reads])
    check(parts, STREAMS[name]["streams"]["qual"])


@pytest.mark.parametrize("name", ["ont_mem", "clr_ratio", "hifi", "multi_ont"])
def test_header_parts_equal_the_reference(name):
    s = synth_case(name)
    want = STREAMS[name]["streams"]["header"]
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        parts = ctx.xhdr_encode(list(s.headers), None, [md for md, _, _ in want])
    check(parts, want)


def test_three_packs_models_live_on():
    """Three read packs: the coder restarts per pack, the models do not (entr_read.h:56-80, entr_qual.h:100-126)."""
    name = "multi_ont"
    s = synth_case(name)
    want = STREAMS[name]["streams"]
    packs = [md for md, _, _ in want["dna"]]
    p = ONT_MEM
    with lib.Context(p["k"], p["modulo"], p["min_count"], p["max_count"], p["max_candidates"]) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        st = ctx.count_finalize()
        mean_len = int(st["tot_kmers"] * p["modulo"] / max(1, s.n_reads) + p["k"] - 1)
        rng = max(1, int(st["n_unique_counted"] * p["modulo"] / max(1, mean_len)))
        ctx.graph_build(lib.sampler(rng, 1.0, 0, s.n_reads))
        ctx.encode(p, np.array(packs, np.uint32))
        check(ctx.xdna_encode(1, packs), want["dna"])
        check(ctx.xqual_encode("4-avg", 0, 1, [7, 14, 26], s.quals, s.offsets, packs), want["qual"])


@pytest.mark.parametrize("seed,n_frac", [(5, 0.0), (6, 0.2)])
def test_random_inputs_equal_the_cpu_restatement(seed, n_frac):
    """Inputs the goldens do not hold (reads with N, ragged packs, an empty pack): device parts = oracle/stage3_exact.c."""
    s = synth.generate(500, 60000, 1500, seed=seed, profile="ont", n_frac=n_frac)
    packs = [137, 0, 300, 63]
    P = dict(anchor_len=16, k=20, modulo=12, hifi=0, min_part_len_alt=64, max_recurence=3, min_anchors=1, min_mmer_frac=0.5, min_mmer_force=0.9, max_matches_mult=10.0, es_cost_mult=1.0)
    headers = [h + (b" x" * (i % 3)) for i, h in enumerate(s.headers)]
    plus = (np.arange(s.n_reads) % 7 == 0).astype(np.uint8)
    with lib.Context(20, 12, 3, 80, 5) as ctx:
        ctx.append_reads(s.bases, s.offsets)
        ctx.count_finalize()
        sampled = np.ones(s.n_reads, np.uint8)
        ctx.graph_build(sampled)
        ctx.encode(P, np.array([p for p in packs if p], np.uint32))
        off, es = ctx.encoded(s.n_reads)
        es_list = [es[int(off[i]):int(off[i + 1])].tobytes() for i in range(s.n_reads)]
        has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(s.n_reads)], np.uint8)
        is_ref = (sampled & (1 - has_n)).astype(np.uint8)
        for level in (1, 2, 3):
            assert ctx.xdna_encode(level, packs) == oracle_lib.xdna_encode(level, 5, es_list, s.bases, s.offsets, is_ref, packs)
            for mode, thr in (("org", []), ("4-avg", [7, 14, 26]), ("2-fix", [7]), ("5-fix", [7, 14, 26, 93])):
                assert ctx.xqual_encode(mode, 0, level, thr, s.quals, s.offsets, packs) == oracle_lib.xqual_encode(mode, 0, level, thr, s.bases, s.quals, s.offsets, packs, es_list)
        assert ctx.xhdr_encode(headers, plus, packs) == oracle_lib.xhdr_encode(headers, plus, packs)
