"""Multi-GPU sharding with the global reference-read set, checked on one GPU: a shard that receives the reference
reads of the earlier shards as context reads produces, for its own reads, exactly the tuples and candidates the single-GPU run
produces for them; its streams decode to its reads.  GPU only."""
import numpy as np
import pytest

import oracle_lib
from colord_b200 import lib, synth

pytestmark = pytest.mark.gpu

P_BAL = dict(anchor_len=16, k=20, modulo=9, hifi=0, min_part_len_alt=48, max_recurence=5, min_anchors=1,
             min_mmer_frac=0.5, min_mmer_force=0.9, max_matches_mult=10.0, es_cost_mult=1.0)


def _slice(s, lo, hi):
    o = s.offsets
    return s.bases[int(o[lo]):int(o[hi])], s.quals[int(o[lo]):int(o[hi])], (o[lo:hi + 1] - o[lo]).astype(np.uint64)


@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("level", [1, 2])
def test_shard_with_context_reads_equals_single_gpu(sparse, level):
    s = synth.generate(900, 200000, 3000, seed=77, profile="ont", n_frac=0.02)
    n, k, f, lo_c, hi_c, c = s.n_reads, 20, 9, 3, 100, 8
    packs = [300, 300, 300]
    sampled = lib.sampler(60, 1.0, 0, n) if sparse else np.ones(n, np.uint8)
    with lib.Context(k, f, lo_c, hi_c, c) as A:
        A.append_reads(s.bases, s.offsets)
        stats = A.count_finalize()
        km, ct = A.filter_list()
        has_n = A.reads_have_n(n)
        A.graph_build(sampled)
        candA, cnA = A.graph_candidates()
        A.encode(P_BAL, packs)
        offA, esA = A.encoded(n)
        for shard in (1, 2):
            lo, hi = 300 * shard, 300 * (shard + 1)
            ids = np.nonzero((sampled[:lo] != 0) & (has_n[:lo] == 0))[0].astype(np.uint32)
            lens = (s.offsets[ids + 1] - s.offsets[ids]).astype(np.uint64)
            ctx_off = np.zeros(len(ids) + 1, np.uint64); ctx_off[1:] = np.cumsum(lens)
            ctx_bases = A.reads_export(ids, int(ctx_off[-1]))           # what the earlier ranks would send
            assert np.array_equal(ctx_bases, np.concatenate([s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] for i in ids]))
            b, q, o = _slice(s, lo, hi)
            with lib.Context(k, f, lo_c, hi_c, c) as B:
                B.append_reads(b, o)
                B.count_finalize()
                B.filter_import(km, ct, stats)
                B.append_context_reads(ctx_bases, ctx_off)
                nc = len(ids)
                B.graph_build(sampled[lo:hi])
                candB, cnB = B.graph_candidates()
                assert np.array_equal(cnB[nc:], cnA[lo:hi])
                valid = np.arange(c)[None, :] < cnA[lo:hi, None]           # slots beyond cand_n are not defined
                assert np.array_equal(candB[nc:][valid], candA[lo:hi][valid])
                assert not cnB[:nc].any()
                B.encode(P_BAL, [300])
                offB, esB = B.encoded(nc + 300)
                assert np.array_equal(offB[nc:] - offB[nc], offA[lo:hi + 1] - offA[lo])
                assert np.array_equal(esB, esA[int(offA[lo]):int(offA[hi])])       # tuples of the shard: bit-identical
                B.dna_encode(level, [300])
                stream, _ = B.dna_stream()
                is_ref = ((sampled[lo:hi] != 0) & (has_n[lo:hi] == 0)).astype(np.uint8)
                dec, doff = oracle_lib.dna_decode(stream, 300, is_ref, len(b), ctx_bases, ctx_off)
                assert np.array_equal(doff, o) and np.array_equal(dec, b)
                B.qual_encode(4, [7, 14, 26], level, q, o, [300])
                qs = B.qual_stream()
                es_sh = esA[int(offA[lo]):int(offA[hi])] if level > 1 else None
                off_sh = (offA[lo:hi + 1] - offA[lo]) if level > 1 else None
                want = oracle_lib.qual_encode(oracle_lib.qual_params(4, [7, 14, 26], level), b, q, o, [300], es_sh, off_sh)
                assert np.array_equal(qs, want)


def test_context_reads_call_order():
    s = synth.generate(50, 20000, 1000, seed=5, profile="ont")
    with lib.Context(20, 9, 3, 100, 8) as B:
        B.append_reads(s.bases, s.offsets)
        with pytest.raises(lib.ClbError):           # before clb_count_finalize
            B.append_context_reads(s.bases[:1000], np.array([0, 1000], np.uint64))
        B.count_finalize()
        B.append_context_reads(s.bases[:1000], np.array([0, 1000], np.uint64))
        with pytest.raises(lib.ClbError):           # only once
            B.append_context_reads(s.bases[:1000], np.array([0, 1000], np.uint64))
