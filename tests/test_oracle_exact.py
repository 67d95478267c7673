"""Pin oracle/stage3_exact.c — the reference's three entropy coders restated on the CPU — against archive parts written by the
UNMODIFIED reference CLI (tests/golden/streams.json: metadata, size and SHA-1 of every part of the "dna", "qual" and "header"
streams; generator tests/golden/make_stream_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

import golden_io
import oracle_lib
from colord_b200 import synth

with open(os.path.join(golden_io.GOLDEN, "streams.json")) as f:
    STREAMS = json.load(f)

# -q mode, forward thresholds (arg_parse.cpp:28-84), data source, level of every golden case
QUAL = {"ont_mem": ("4-avg", [7, 14, 26], 0, 1), "ont_bal": ("4-avg", [7, 14, 26], 0, 2), "clr_ratio": ("none", [], 1, 3), "hifi": ("5-avg", [7, 14, 26, 93], 2, 2),
        "q_org": ("org", [], 0, 1), "q_org_bal": ("org", [], 0, 2), "q_2avg": ("2-avg", [7], 0, 1), "q_5avg": ("5-avg", [7, 14, 26, 93], 0, 1), "q_2fix": ("2-fix", [7], 0, 1),
        "q_4fix": ("4-fix", [7, 14, 26], 0, 1), "q_5fix": ("5-fix", [7, 14, 26, 93], 0, 1), "q_avg": ("avg", [], 0, 1), "q_none": ("none", [], 0, 1),
        "q_4fix_thr": ("4-fix", [5, 12, 20], 0, 1), "q_hifi_org": ("org", [], 2, 2), "q_clr_org": ("org", [], 1, 1)}


def check(parts, want, metadata=None):
    assert len(parts) == len(want)
    for i, (p, (md, size, sha)) in enumerate(zip(parts, want)):
        assert len(p) == size, (i, len(p), size)
        assert hashlib.sha1(p).hexdigest() == sha, i
        if metadata is not None:
            assert md == metadata[i]


def synth_case(name):
    s = synth.generate(**STREAMS[name]["generator"])
    assert hashlib.sha1(s.bases.tobytes()).hexdigest() == STREAMS[name]["bases_sha1"], "synthetic input drifted"
    return s


@pytest.mark.parametrize("name", ["ont_mem", "ont_bal", "clr_ratio", "hifi"])
def test_dna_stream_equals_reference(name):
    """CompactES tuples as the reference's CEncoder emitted them -> the bytes of the reference's `dna` parts."""
    g = golden_io.load_case(name)
    s = g.reads_in
    parts = oracle_lib.xdna_encode(g.params["level"], g.params["max_candidates"], g.es, s.bases, s.offsets, g.is_ref, g.es_packs)
    check(parts, STREAMS[name]["streams"]["dna"], metadata=g.es_packs)


@pytest.mark.parametrize("name", list(QUAL))
def test_quality_stream_equals_reference(name):
    mode, thr, source, level = QUAL[name]
    if level > 1:      # the per-base match / anchor flags come from the tuples: the reference's own CompactES dump of the same input
        g = golden_io.load_case({"q_hifi_org": "hifi", "q_org_bal": "ont_bal"}.get(name, name))
        s, es = g.reads_in, g.es
    else:
        s, es = synth_case(name), None
    parts = oracle_lib.xqual_encode(mode, source, level, thr, s.bases, s.quals, s.offsets, [s.n_reads], es)
    check(parts, STREAMS[name]["streams"]["qual"])


@pytest.mark.parametrize("name", ["ont_mem", "clr_ratio", "hifi", "multi_ont"])
def test_header_stream_equals_reference(name):
    s = synth_case(name)
    want = STREAMS[name]["streams"]["header"]
    parts = oracle_lib.xhdr_encode(list(s.headers), np.zeros(s.n_reads, np.uint8), [md for md, _, _ in want])
    check(parts, want)


# compress-ont defaults for inputs below 4e9 bases estimated size (arg_parse.cpp:89-408, compression.cpp:41-93); golden params.txt keys
ONT_MEM = dict(k=20, anchor_len=16, modulo=12, min_count=4, max_count=80, max_candidates=5, level=1, sparse=1, hifi=0, min_part_len_alt=64, max_recurence=3,
               min_anchors=1, es_cost_mult=1.0, min_mmer_frac=0.5, min_mmer_force=0.9, max_matches_mult=10.0)


def test_whole_pipeline_three_packs_equals_reference():
    """Oracle stages 1 + 2 + 3 on an input of three read packs against the stock binary's archive: the coder restarts per pack while
    its models live on (entr_read.h:56-80, entr_qual.h:100-126), the qual parts follow the dna parts."""
    name = "multi_ont"
    s = synth_case(name)
    es, is_ref, packs = oracle_lib.pipeline(s, ONT_MEM)
    want = STREAMS[name]["streams"]
    assert packs == [md for md, _, _ in want["dna"]]
    check(oracle_lib.xdna_encode(1, 5, es, s.bases, s.offsets, is_ref, packs), want["dna"])
    check(oracle_lib.xqual_encode("4-avg", 0, 1, [7, 14, 26], s.bases, s.quals, s.offsets, packs), want["qual"])
