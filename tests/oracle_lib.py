"""ctypes binding to oracle/liboracle.so (the plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)


def lib():
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    L.orc_murmur64.restype = C.c_uint64
    L.orc_murmur64.argtypes = [C.c_uint64]
    L.orc_count_kmers.restype = C.c_uint64
    L.orc_count_kmers.argtypes = [_u8p, _u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  _u64p, _u32p, C.c_uint64, _u64p]
    L.orc_accepted_kmers.restype = C.c_uint64
    L.orc_accepted_kmers.argtypes = [_u8p, _u64p, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, C.c_uint64,
                                     _u64p, _u64p, C.c_uint64]
    L.orc_sampler.restype = None
    L.orc_sampler.argtypes = [C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, _u8p]
    L.orc_sim_graph.restype = C.c_uint64
    L.orc_sim_graph.argtypes = [_u64p, _u64p, C.c_uint32, _u8p, _u8p, C.c_uint32, C.c_uint32, _u32p, _u32p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.orc_pack_ref_read.restype = C.c_uint64
    L.orc_pack_ref_read.argtypes = [_u8p, C.c_uint64, _u8p]
    return L


def count_kmers(bases, offsets, k, modulo, min_count, max_count):
    L = lib()
    n = len(offsets) - 1
    stats = np.zeros(5, np.uint64)
    cap = max(1024, int(offsets[-1]) // max(1, modulo) + 1024)
    km = np.zeros(cap, np.uint64)
    ct = np.zeros(cap, np.uint32)
    ns = L.orc_count_kmers(bases, offsets, n, k, modulo, min_count, max_count, km, ct, cap, stats)
    assert ns <= cap
    return km[:ns].copy(), ct[:ns].copy(), dict(n_reads=int(stats[0]), tot_kmers=int(stats[1]), n_unique=int(stats[2]),
                                                 n_unique_counted=int(stats[3]), total_count_filtered=int(stats[4]))


def accepted_kmers(bases, offsets, k, modulo, kmer_set_sorted):
    L = lib()
    n = len(offsets) - 1
    cap = max(1024, int(offsets[-1]) // max(1, modulo) * 2 + 1024)
    off = np.zeros(n + 1, np.uint64)
    acc = np.zeros(cap, np.uint64)
    tot = L.orc_accepted_kmers(bases, offsets, n, k, modulo, np.ascontiguousarray(kmer_set_sorted), len(kmer_set_sorted), off, acc, cap)
    assert tot <= cap
    return off, acc[:tot].copy()


def sampler(rng_range, exponent, n_pseudo, n):
    L = lib()
    out = np.zeros(n, np.uint8)
    L.orc_sampler(rng_range, exponent, n_pseudo, n, out)
    return out


def sim_graph(acc_off, acc, has_n, sampled, max_candidates, max_kmer_count, hifi=False):
    L = lib()
    n = len(acc_off) - 1
    cand = np.zeros(n * max_candidates, np.uint32)
    cand_n = np.zeros(n, np.uint32)
    if hifi:
        common_off = np.zeros(n * max_candidates, np.uint64)
        common_n = np.zeros(n * max_candidates, np.uint32)
        cap = max(1024, 4 * len(acc) * 1 + 1024)
        common = np.zeros(cap, np.uint64)
        tot = L.orc_sim_graph(acc_off, np.ascontiguousarray(acc), n, has_n, sampled, max_candidates, max_kmer_count, cand, cand_n,
                              common_off.ctypes.data, common_n.ctypes.data, common.ctypes.data, cap)
        assert tot <= cap
        return cand.reshape(n, max_candidates), cand_n, (common_off, common_n, common[:tot])
    L.orc_sim_graph(acc_off, np.ascontiguousarray(acc), n, has_n, sampled, max_candidates, max_kmer_count, cand, cand_n,
                    None, None, None, 0)
    return cand.reshape(n, max_candidates), cand_n, None


class S2Params(C.Structure):
    _fields_ = [("anchor_len", C.c_uint32), ("kmer_len", C.c_uint32), ("modulo", C.c_uint32), ("is_hifi", C.c_uint32),
                ("min_part_len_alt", C.c_uint32), ("max_recurence", C.c_uint32), ("min_anchors", C.c_uint32),
                ("min_mmer_frac", C.c_double), ("min_mmer_force", C.c_double), ("max_matches_mult", C.c_double), ("es_cost_mult", C.c_double)]


def s2_params(p):
    """From a golden params.txt dict (or the same keys)."""
    return S2Params(p["anchor_len"], p["k"], p["modulo"], int(p.get("hifi", 0)), p["min_part_len_alt"], p["max_recurence"], p["min_anchors"],
                    float(p["min_mmer_frac"]), float(p["min_mmer_force"]), float(p["max_matches_mult"]), float(p["es_cost_mult"]))


def edit_script(ref_sym, enc_sym, kind, ref_tail=b"\xff", enc_tail=b"\xff"):
    """ref_sym/enc_sym: uint8 arrays of symbols 0..3 (the part); *_tail: what follows the part in the read (guard)."""
    L = lib()
    L.orc_edit_script.restype = C.c_uint32
    L.orc_edit_script.argtypes = [_u8p, C.c_uint32, _u8p, C.c_uint32, C.c_int, C.c_char_p, C.c_uint32]
    r = np.concatenate([np.asarray(ref_sym, np.uint8), np.frombuffer(ref_tail, np.uint8)])
    e = np.concatenate([np.asarray(enc_sym, np.uint8), np.frombuffer(enc_tail, np.uint8)])
    cap = 2 * (len(r) + len(e)) + 16
    buf = C.create_string_buffer(cap)
    n = L.orc_edit_script(np.ascontiguousarray(r), len(ref_sym), np.ascontiguousarray(e), len(enc_sym), kind, buf, cap)
    assert n <= cap
    return buf.raw[:n]


def encode_reads(bases, offsets, is_ref, cand, cand_n, pack_sizes, params, common=None):
    """-> list of CompactES byte strings, one per read (oracle/stage2.c: orc_encode_reads)."""
    L = lib()
    L.orc_encode_reads.restype = C.c_uint64
    L.orc_encode_reads.argtypes = [_u8p, _u64p, C.c_uint32, _u8p, _u32p, _u32p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   _u32p, C.c_uint32, C.POINTER(S2Params), _u8p, C.c_uint64, _u64p]
    n = len(offsets) - 1
    mc = cand.shape[1]
    cand = np.ascontiguousarray(cand, np.uint32).reshape(-1)
    cand_n = np.ascontiguousarray(cand_n, np.uint32)
    pack_sizes = np.ascontiguousarray(pack_sizes, np.uint32)
    cap = int(offsets[-1]) * 2 + 64 * n + 1024
    out = np.zeros(cap, np.uint8)
    es_off = np.zeros(n + 1, np.uint64)
    if common is not None:
        coff, cn, ckm = [np.ascontiguousarray(x) for x in common]
        if len(ckm) == 0:
            ckm = np.zeros(1, np.uint64)
        args = (coff.ctypes.data, cn.ctypes.data, ckm.ctypes.data)
    else:
        args = (None, None, None)
    prm = params if isinstance(params, S2Params) else s2_params(params)
    tot = L.orc_encode_reads(np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(offsets, np.uint64), n, np.ascontiguousarray(is_ref, np.uint8),
                             cand, cand_n, mc, *args, pack_sizes, len(pack_sizes), C.byref(prm), out, cap, es_off)
    assert tot <= cap
    return [out[int(es_off[i]):int(es_off[i + 1])].tobytes() for i in range(n)]


def candidates(bases, offsets, is_ref, cand, cand_n, params):
    """-> per read a list of (ref_id, rev, tot, [(len, pos_enc, pos_ref), ...]) (oracle/stage2.c: orc_candidates)."""
    L = lib()
    L.orc_candidates.restype = C.c_uint64
    L.orc_candidates.argtypes = [_u8p, _u64p, C.c_uint32, _u8p, _u32p, _u32p, C.c_uint32, C.POINTER(S2Params), _u64p, _u32p, C.c_uint64]
    n = len(offsets) - 1
    mc = cand.shape[1]
    cand = np.ascontiguousarray(cand, np.uint32).reshape(-1)
    cand_n = np.ascontiguousarray(cand_n, np.uint32)
    prm = params if isinstance(params, S2Params) else s2_params(params)
    off = np.zeros(n + 1, np.uint64)
    cap = int(offsets[-1]) * mc + 64
    data = np.zeros(cap, np.uint32)
    tot = L.orc_candidates(np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(offsets, np.uint64), n, np.ascontiguousarray(is_ref, np.uint8),
                           cand, cand_n, mc, C.byref(prm), off, data, cap)
    assert tot <= cap
    out = []
    for i in range(n):
        rec, p, e = [], int(off[i]), int(off[i + 1])
        while p < e:
            ref_id, rev, t, na = (int(x) for x in data[p:p + 4]); p += 4
            rec.append((ref_id, rev, t, [tuple(int(x) for x in data[p + 3 * k:p + 3 * k + 3]) for k in range(na)]))
            p += 3 * na
        out.append(rec)
    return out


# ---------------------------------------------------------------------------------------------------- stage 3: quality
class QualParams(C.Structure):
    _fields_ = [("n_bins", C.c_uint32), ("thr", C.c_uint32 * 4), ("level", C.c_uint32)]


def qual_params(n_bins, thr, level):
    p = QualParams()
    p.n_bins, p.level = n_bins, level
    for i, t in enumerate(thr):
        p.thr[i] = t
    return p


def qual_lossy(P, bases, quals, offsets):
    """The reference's lossy quality transform (what `colord decompress` prints) — oracle/stage3_qual.c."""
    L = lib()
    L.orc_qual_lossy.restype = None
    L.orc_qual_lossy.argtypes = [C.POINTER(QualParams), _u8p, _u8p, _u64p, C.c_uint32, _u8p]
    out = np.zeros(len(quals), np.uint8)
    L.orc_qual_lossy(C.byref(P), np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(quals, np.uint8), np.ascontiguousarray(offsets, np.uint64), len(offsets) - 1, out)
    return out


def qual_encode(P, bases, quals, offsets, pack_sizes, es=None, es_off=None):
    L = lib()
    L.orc_qual_encode.restype = C.c_uint64
    L.orc_qual_encode.argtypes = [C.POINTER(QualParams), _u8p, _u8p, _u64p, C.c_uint32, C.c_void_p, C.c_void_p, _u32p, C.c_uint32, _u8p, C.c_uint64]
    cap = int(len(quals)) + (1 << 20) + (8 << (P.level > 1 and 21 or 19))
    out = np.zeros(cap, np.uint8)
    ps = np.ascontiguousarray(pack_sizes, np.uint32)
    esp = es.ctypes.data if es is not None else None
    eop = es_off.ctypes.data if es_off is not None else None
    n = L.orc_qual_encode(C.byref(P), np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(quals, np.uint8), np.ascontiguousarray(offsets, np.uint64),
                          len(offsets) - 1, esp, eop, ps, len(ps), out, cap)
    assert n <= cap
    return out[:n].copy()


def qual_decode(stream, bases, offsets, es=None, es_off=None):
    L = lib()
    L.orc_qual_decode.restype = C.c_int
    L.orc_qual_decode.argtypes = [_u8p, C.c_uint64, _u8p, _u64p, C.c_uint32, C.c_void_p, C.c_void_p, _u8p]
    out = np.zeros(int(offsets[-1]), np.uint8)
    esp = es.ctypes.data if es is not None else None
    eop = es_off.ctypes.data if es_off is not None else None
    rc = L.orc_qual_decode(np.ascontiguousarray(stream, np.uint8), len(stream), np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(offsets, np.uint64),
                           len(offsets) - 1, esp, eop, out)
    assert rc == 0, rc
    return out


def dna_decode(stream, n_reads, is_ref, cap_bases, ctx_bases=None, ctx_off=None):
    """Decoder of the native DNA container (oracle/stage3_dna.c) -> (bases ASCII, offsets).  ctx_bases / ctx_off: the context
    reads of a shard's container (reference reads of earlier shards)."""
    L = lib()
    L.orc_dna_decode_ctx.restype = C.c_int64
    L.orc_dna_decode_ctx.argtypes = [_u8p, C.c_uint64, C.c_uint32, _u8p, _u8p, _u64p, C.c_uint32, _u8p, C.c_uint64, _u64p]
    out = np.zeros(int(cap_bases) + 16, np.uint8)
    off = np.zeros(n_reads + 1, np.uint64)
    if ctx_off is None:
        ctx_bases, ctx_off = np.zeros(1, np.uint8), np.zeros(1, np.uint64)
    rc = L.orc_dna_decode_ctx(np.ascontiguousarray(stream, np.uint8), len(stream), n_reads, np.ascontiguousarray(is_ref if len(is_ref) else [0], np.uint8),
                              np.ascontiguousarray(ctx_bases if len(ctx_bases) else [0], np.uint8), np.ascontiguousarray(ctx_off, np.uint64), len(ctx_off) - 1,
                              out, int(cap_bases), off)
    assert rc == 0, rc
    return out[:int(off[-1])], off


def _hdr_arrays(headers):
    b = np.frombuffer(b"".join(headers), np.uint8)
    off = np.zeros(len(headers) + 1, np.uint64)
    off[1:] = np.cumsum([len(h) for h in headers], dtype=np.uint64)
    return np.ascontiguousarray(b if len(b) else np.zeros(1, np.uint8)), off


def hdr_encode(headers, plus=None, packs=None):
    """CPU twin of the native header container (oracle/stage3_hdr.c).  packs None: packs of 4096 headers (the device default)."""
    L = lib()
    L.orc_hdr_encode.restype = C.c_int64
    L.orc_hdr_encode.argtypes = [_u8p, _u64p, C.c_void_p, C.c_uint64, _u32p, C.c_uint32, _u8p, C.c_uint64]
    n = len(headers)
    if packs is None:
        packs = [min(4096, n - i) for i in range(0, n, 4096)]
    ps = np.ascontiguousarray(packs if len(packs) else [0], np.uint32)
    b, off = _hdr_arrays(headers)
    pl = None if plus is None else np.ascontiguousarray(plus, np.uint8)
    cap = 2 * len(b) + 300 * (len(packs) + 1) + (1 << 20)
    out = np.zeros(cap, np.uint8)
    rc = L.orc_hdr_encode(b, off, None if pl is None else pl.ctypes.data, n, ps, len(packs), out, cap)
    assert rc >= 0, rc
    return out[:rc]


def hdr_decode(stream, n, cap_bytes):
    """Decoder of the native header container -> (list of bytes, plus flags)."""
    L = lib()
    L.orc_hdr_decode.restype = C.c_int64
    L.orc_hdr_decode.argtypes = [_u8p, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p, _u8p]
    out = np.zeros(int(cap_bytes) + 16, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    plus = np.zeros(n + 1, np.uint8)
    rc = L.orc_hdr_decode(np.ascontiguousarray(stream, np.uint8), len(stream), n, out, int(cap_bytes), off, plus)
    assert rc >= 0, rc
    raw = out.tobytes()
    return [raw[int(off[i]):int(off[i + 1])] for i in range(n)], plus[:n]


def qorg_encode(source, level, bases, quals, offsets, pack_sizes, es=None, es_off=None):
    """CPU twin of the native lossless-quality container (oracle/stage3_qorg.c)."""
    L = lib()
    L.orc_qorg_encode.restype = C.c_int64
    L.orc_qorg_encode.argtypes = [C.c_uint32, C.c_uint32, _u8p, _u8p, _u64p, C.c_uint32, C.c_void_p, C.c_void_p, _u32p, C.c_uint32, _u8p, C.c_uint64]
    cap = int(len(quals)) + (64 << 20)
    out = np.zeros(cap, np.uint8)
    ps = np.ascontiguousarray(pack_sizes, np.uint32)
    esp = es.ctypes.data if es is not None else None
    eop = es_off.ctypes.data if es_off is not None else None
    n = L.orc_qorg_encode(source, level, np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(quals, np.uint8), np.ascontiguousarray(offsets, np.uint64),
                          len(offsets) - 1, esp, eop, ps, len(ps), out, cap)
    assert 0 <= n <= cap, n
    return out[:n].copy()


def qorg_decode(stream, bases, offsets, es=None, es_off=None):
    L = lib()
    L.orc_qorg_decode.restype = C.c_int
    L.orc_qorg_decode.argtypes = [_u8p, C.c_uint64, _u8p, _u64p, C.c_uint32, C.c_void_p, C.c_void_p, _u8p]
    out = np.zeros(int(offsets[-1]) + 1, np.uint8)
    esp = es.ctypes.data if es is not None else None
    eop = es_off.ctypes.data if es_off is not None else None
    rc = L.orc_qorg_decode(np.ascontiguousarray(stream, np.uint8), len(stream), np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(offsets, np.uint64),
                           len(offsets) - 1, esp, eop, out)
    assert rc == 0, rc
    return out[:int(offsets[-1])]


def dna_encode(level, max_cand, es_list, bases, offsets, is_ref, pack_sizes):
    """CPU twin of the native DNA container's encoder (oracle/stage3_dna.c: orc_dna_encode) from per-read CompactES byte strings."""
    L = lib()
    L.orc_dna_encode.restype = C.c_int64
    L.orc_dna_encode.argtypes = [C.c_uint32, C.c_uint32, _u8p, _u64p, _u8p, _u64p, _u8p, C.c_uint32, _u32p, C.c_uint32, _u8p, C.c_uint64]
    es = np.frombuffer(b"".join(es_list), np.uint8).copy()
    es_off = np.zeros(len(es_list) + 1, np.uint64)
    es_off[1:] = np.cumsum([len(x) for x in es_list], dtype=np.uint64)
    cap = int(len(es)) + int(len(bases)) // 2 + (8 << 20)
    out = np.zeros(cap, np.uint8)
    ps = np.ascontiguousarray(pack_sizes, np.uint32)
    n = L.orc_dna_encode(level, max_cand, es, es_off, np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(offsets, np.uint64),
                         np.ascontiguousarray(is_ref, np.uint8), len(es_list), ps, len(ps), out, cap)
    assert 0 <= n <= cap, n
    return out[:n].copy()


# ---------------------------------------------------------------------------------------------- exact (reference-format) streams
def _split_parts(out, sizes):
    parts, at = [], 0
    for s in sizes:
        parts.append(out[at:at + int(s)].tobytes())
        at += int(s)
    return parts


def _es_arrays(es_list):
    n = len(es_list)
    es_off = np.zeros(n + 1, np.uint64)
    es_off[1:] = np.cumsum([len(e) for e in es_list])
    es = np.frombuffer(b"".join(es_list), np.uint8).copy() if n and es_off[-1] else np.zeros(1, np.uint8)
    return es, es_off


def xdna_encode(level, max_cand, es_list, bases, offsets, is_ref, pack_sizes, n_skip=0):
    """oracle/stage3_exact.c: the reference's `dna` stream parts (list of bytes, one per pack)."""
    L = lib()
    L.orc_xdna_encode.restype = C.c_int64
    L.orc_xdna_encode.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u64p, _u8p, _u64p, _u8p, C.c_uint32, _u32p, C.c_uint32, _u8p, C.c_uint64, _u64p]
    es, es_off = _es_arrays(es_list)
    ps = np.ascontiguousarray(pack_sizes, np.uint32)
    sizes = np.zeros(len(ps) + 1, np.uint64)
    cap = len(es) + 64 * len(ps) + 1024
    out = np.zeros(cap, np.uint8)
    n = L.orc_xdna_encode(level, max_cand, n_skip, es, es_off, np.ascontiguousarray(bases), np.ascontiguousarray(offsets, np.uint64),
                          np.ascontiguousarray(is_ref, np.uint8), len(offsets) - 1, ps, len(ps), out, cap, sizes)
    assert n >= 0
    return _split_parts(out, sizes[:len(ps)])


QMODES = {"org": 0, "5-avg": 1, "4-avg": 2, "2-avg": 3, "5-fix": 4, "4-fix": 5, "2-fix": 6, "avg": 7, "none": 8}


def xqual_encode(mode, source, level, thr, bases, quals, offsets, pack_sizes, es_list=None):
    """oracle/stage3_exact.c: the reference's `qual` stream parts.  mode: a -q name; source 0 ONT / 1 CLR / 2 HiFi."""
    L = lib()
    L.orc_xqual_encode.restype = C.c_int64
    L.orc_xqual_encode.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _u32p, _u8p, _u8p, _u64p, _u8p, _u64p, C.c_uint32, _u32p, C.c_uint32, _u8p, C.c_uint64, _u64p]
    n = len(offsets) - 1
    es, es_off = _es_arrays(es_list if es_list is not None else [bytes([9 << 4])] * n)
    ps = np.ascontiguousarray(pack_sizes, np.uint32)
    sizes = np.zeros(len(ps) + 1, np.uint64)
    cap = int(offsets[-1]) * 2 + 64 * len(ps) + 4096
    out = np.zeros(cap, np.uint8)
    t = np.zeros(8, np.uint32)
    t[:len(thr)] = thr
    r = L.orc_xqual_encode(QMODES[mode], source, level, t, np.ascontiguousarray(bases), np.ascontiguousarray(quals), np.ascontiguousarray(offsets, np.uint64),
                           es, es_off, n, ps, len(ps), out, cap, sizes)
    assert r >= 0
    return _split_parts(out, sizes[:len(ps)])


def xhdr_encode(headers, plus, pack_sizes):
    """oracle/stage3_exact.c: the reference's `header` stream parts.  headers: list of bytes without the leading '@' / '>'."""
    L = lib()
    L.orc_xhdr_encode.restype = C.c_int64
    L.orc_xhdr_encode.argtypes = [_u8p, _u64p, _u8p, C.c_uint32, _u32p, C.c_uint32, _u8p, C.c_uint64, _u64p]
    n = len(headers)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum([len(h) for h in headers])
    by = np.frombuffer(b"".join(headers), np.uint8).copy() if n and off[-1] else np.zeros(1, np.uint8)
    ps = np.ascontiguousarray(pack_sizes, np.uint32)
    sizes = np.zeros(len(ps) + 1, np.uint64)
    cap = len(by) * 2 + 64 * len(ps) + 4096
    out = np.zeros(cap, np.uint8)
    r = L.orc_xhdr_encode(by, off, np.ascontiguousarray(plus, np.uint8), n, ps, len(ps), out, cap, sizes)
    assert r >= 0
    return _split_parts(out, sizes[:len(ps)])


def read_packs(offsets, limit=2 << 21):
    """Pack sizes of the reader (in_reads.cpp:62-76: a pack closes once its read_t bytes — length + guard — reach 4 MiB)."""
    packs, cur, n = [], 0, 0
    for ln in np.diff(np.asarray(offsets, np.int64)):
        cur += int(ln) + 1
        n += 1
        if cur >= limit:
            packs.append(n)
            cur = n = 0
    if n:
        packs.append(n)
    return packs


def pipeline(s, p):
    """Stages 1 + 2 of the oracle on synthetic reads `s` with the parameter dict `p` (keys as in a golden params.txt).
    -> (CompactES list, is_ref, pack sizes)"""
    km, ct, st = count_kmers(s.bases, s.offsets, p["k"], p["modulo"], p["min_count"], p["max_count"])
    off, acc = accepted_kmers(s.bases, s.offsets, p["k"], p["modulo"], km)
    if p.get("sparse", 1):
        mean_len = int(st["tot_kmers"] * p["modulo"] / max(1, s.n_reads) + p["k"] - 1)
        rng = max(1, int(p.get("sparse_g", 1.0) * st["n_unique_counted"] * p["modulo"] / max(1, mean_len)))
        sampled = sampler(rng, p.get("sparse_exponent", 1.0), 0, s.n_reads)
    else:
        sampled = np.ones(s.n_reads, np.uint8)
    has_n = np.array([(s.bases[int(s.offsets[i]):int(s.offsets[i + 1])] == ord("N")).any() for i in range(s.n_reads)], np.uint8)
    cand, cn, common = sim_graph(off, acc, has_n, sampled, p["max_candidates"], p["max_count"], hifi=bool(p.get("hifi", 0)))
    is_ref = (sampled & (1 - has_n)).astype(np.uint8)
    packs = read_packs(s.offsets)
    es = encode_reads(s.bases, s.offsets, is_ref, cand, cn, np.array(packs, np.uint32), s2_params(p), common)
    return es, is_ref, packs
