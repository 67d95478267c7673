"""ctypes binding of include/colord_b200.h.  Loading fails loudly if the CUDA library was not built."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CLB_LIBRARY") or os.path.join(_HERE, "libcolord_b200.so")      # CLB_LIBRARY: a profiling build (make phases)

STATUS = {0: "OK", 1: "NO_DEVICE", 2: "CUDA", 3: "BAD_ARG", 4: "BAD_SYMBOL", 5: "STATE", 6: "CAPACITY"}
EXPORTS = [
    "clb_create", "clb_destroy", "clb_last_error", "clb_set_stream", "clb_synchronize", "clb_append_reads",
    "clb_counts_size", "clb_counts_export", "clb_counts_sizes", "clb_counts_export_all", "clb_counts_reset", "clb_counts_merge", "clb_count_finalize",
    "clb_filter_list", "clb_filter_import", "clb_filter_check", "clb_graph_build", "clb_graph_accepted_size",
    "clb_graph_accepted", "clb_graph_candidates", "clb_graph_common_size", "clb_graph_common", "clb_get_packed_read",
    "clb_sampler", "clb_kernel_launches", "clb_profile_enable", "clb_profile_get", "clb_edit_scripts",
    "clb_encode", "clb_encode_size", "clb_encode_get", "clb_encode_keep_candidates", "clb_encode_candidates_size", "clb_encode_candidates",
    "clb_qual_encode", "clb_qual_size", "clb_qual_get", "clb_dna_encode", "clb_dna_size", "clb_dna_get", "clb_hdr_encode", "clb_hdr_size", "clb_hdr_get",
    "clb_append_context_reads", "clb_reads_have_n", "clb_reads_export", "clb_qual_encode_original", "clb_release_cached_memory",
    "clb_append_quals", "clb_encode_stats_enable", "clb_encode_stats_get", "clb_count_sequences", "clb_xplain_encode", "clb_host_alloc", "clb_host_free", "clb_xdna_encode", "clb_xqual_encode", "clb_xhdr_encode", "clb_xstream_size", "clb_xstream_get",
]
KERNEL_CLASSES = ["k_pack", "k_count", "k_tab_misc", "k_finalize", "k_accept", "k_postings", "k_vote", "k_common", "k_misc", "k_align", "k_anchors", "k_encode", "k_decide", "k_estimate", "k_emit", "k_qual", "k_dna", "k_hdr"]


class Params(C.Structure):
    _fields_ = [("kmer_len", C.c_uint32), ("modulo", C.c_uint32), ("min_count", C.c_uint32), ("max_count", C.c_uint32),
                ("max_candidates", C.c_uint32), ("is_hifi", C.c_uint32), ("expected_bases", C.c_uint64), ("device", C.c_int32)]


class KmerStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("tot_kmers", C.c_uint64), ("n_unique", C.c_uint64),
                ("n_unique_counted", C.c_uint64), ("total_count_filtered", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class EncodeParams(C.Structure):
    _fields_ = [("anchor_len", C.c_uint32), ("min_part_len_alt", C.c_uint32), ("max_recurence", C.c_uint32), ("min_anchors", C.c_uint32),
                ("min_mmer_frac", C.c_double), ("min_mmer_force", C.c_double), ("max_matches_mult", C.c_double), ("es_cost_mult", C.c_double)]


class QualParams(C.Structure):
    _fields_ = [("n_bins", C.c_uint32), ("thresholds", C.c_uint32 * 4), ("level", C.c_uint32)]


class ClbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"colord_b200: {STATUS.get(status, status)}: {msg}")
        self.status = status


_lib = None


def load():
    """dlopen the library.  No fallback: a missing build is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C colord_b200/csrc); colord_b200 has no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.clb_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.clb_destroy.argtypes = [vp]; L.clb_destroy.restype = None
    L.clb_last_error.argtypes = [vp]; L.clb_last_error.restype = C.c_char_p
    L.clb_set_stream.argtypes = [vp, vp]
    L.clb_synchronize.argtypes = [vp]
    L.clb_append_reads.argtypes = [vp, vp, vp, u32, i32]
    L.clb_counts_size.argtypes = [vp, u32, u32, C.POINTER(u64)]
    L.clb_counts_export.argtypes = [vp, u32, u32, vp, vp, u64, C.POINTER(u64), i32]
    L.clb_counts_sizes.argtypes = [vp, u32, vp]
    L.clb_counts_export_all.argtypes = [vp, u32, vp, vp, vp, u64]
    L.clb_counts_reset.argtypes = [vp]
    L.clb_counts_merge.argtypes = [vp, vp, vp, u64, u64, i32]
    L.clb_count_finalize.argtypes = [vp, C.POINTER(KmerStats)]
    L.clb_filter_list.argtypes = [vp, vp, vp, u64, C.POINTER(u64), i32]
    L.clb_filter_import.argtypes = [vp, vp, vp, u64, C.POINTER(KmerStats), i32]
    L.clb_filter_check.argtypes = [vp, vp, u64, vp, vp]
    L.clb_graph_build.argtypes = [vp, vp, u32]
    L.clb_graph_accepted_size.argtypes = [vp, C.POINTER(u64)]
    L.clb_graph_accepted.argtypes = [vp, vp, vp, u64]
    L.clb_graph_candidates.argtypes = [vp, vp, vp]
    L.clb_graph_common_size.argtypes = [vp, C.POINTER(u64)]
    L.clb_graph_common.argtypes = [vp, vp, vp, vp, u64]
    L.clb_get_packed_read.argtypes = [vp, u32, vp, u64, C.POINTER(u64)]
    L.clb_sampler.argtypes = [u32, C.c_double, u32, u32, vp]; L.clb_sampler.restype = None
    L.clb_release_cached_memory.argtypes = [i32]
    L.clb_kernel_launches.argtypes = [vp]; L.clb_kernel_launches.restype = u64
    L.clb_edit_scripts.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, u64, vp, vp, u64]
    L.clb_encode.argtypes = [vp, C.POINTER(EncodeParams), vp, u32]
    L.clb_encode_size.argtypes = [vp, C.POINTER(u64)]
    L.clb_encode_get.argtypes = [vp, vp, vp, u64, i32]
    L.clb_encode_keep_candidates.argtypes = [vp, i32]
    L.clb_encode_candidates_size.argtypes = [vp, C.POINTER(u64)]
    L.clb_encode_candidates.argtypes = [vp, vp, vp, u64]
    L.clb_dna_encode.argtypes = [vp, u32, vp, u32]
    L.clb_dna_size.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.clb_dna_get.argtypes = [vp, vp, u64, i32]
    L.clb_append_context_reads.argtypes = [vp, vp, vp, u32, i32]
    L.clb_reads_have_n.argtypes = [vp, vp]
    L.clb_reads_export.argtypes = [vp, vp, u32, vp, u64, i32]
    L.clb_qual_encode_original.argtypes = [vp, u32, u32, vp, vp, i32, vp, u32]
    L.clb_hdr_encode.argtypes = [vp, vp, vp, vp, u64, i32, vp, u32]
    L.clb_hdr_size.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.clb_hdr_get.argtypes = [vp, vp, u64, i32]
    L.clb_qual_encode.argtypes = [vp, C.POINTER(QualParams), vp, vp, i32, vp, u32]
    L.clb_qual_size.argtypes = [vp, C.POINTER(u64)]
    L.clb_qual_get.argtypes = [vp, vp, u64, i32]
    L.clb_append_quals.argtypes = [vp, vp, u64, i32]
    L.clb_count_sequences.argtypes = [vp, vp, vp, u32, i32]
    L.clb_xplain_encode.argtypes = [vp, vp, vp, u32, u32]
    L.clb_host_alloc.argtypes = [u64]; L.clb_host_alloc.restype = vp
    L.clb_host_free.argtypes = [vp]; L.clb_host_free.restype = None
    L.clb_xdna_encode.argtypes = [vp, u32, vp, u32]
    L.clb_xqual_encode.argtypes = [vp, u32, u32, u32, vp, vp, vp, i32, vp, u32]
    L.clb_xhdr_encode.argtypes = [vp, vp, vp, vp, u64, i32, vp, u32]
    L.clb_xstream_size.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u32)]
    L.clb_xstream_get.argtypes = [vp, u32, vp, u64, vp, i32]
    L.clb_profile_enable.argtypes = [vp, i32]
    L.clb_profile_get.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(u64)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("clb_destroy", "clb_sampler"):
            fn.restype = C.c_int
    _lib = L
    return L


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def release_cached_memory(device=0):
    """Hand the device memory cached by finished jobs back to the system (clb_release_cached_memory)."""
    return load().clb_release_cached_memory(device)


def sampler(rng_range, exponent, n_pseudo, n):
    out = np.zeros(n, np.uint8)
    load().clb_sampler(rng_range, float(exponent), n_pseudo, n, _np_ptr(out))
    return out


class Context:
    """One library context (= one GPU).  Mirrors the order of calls runCompression makes (compression.cpp:432-564)."""

    def __init__(self, kmer_len, modulo, min_count, max_count, max_candidates, is_hifi=False, expected_bases=0, device=0):
        self.L = load()
        self.params = Params(kmer_len, modulo, min_count, max_count, max_candidates, int(bool(is_hifi)), int(expected_bases), device)
        h = C.c_void_p()
        st = self.L.clb_create(C.byref(self.params), C.byref(h))
        if st != 0:
            raise ClbError(st, self.L.clb_last_error(None).decode())
        self.h = h
        self.n_reads = 0          # reads in the store (context reads included)
        self.n_context = 0

    def _ck(self, st):
        if st != 0:
            raise ClbError(st, self.L.clb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.clb_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_handle):
        self._ck(self.L.clb_set_stream(self.h, C.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        self._ck(self.L.clb_synchronize(self.h))

    # ---- stage 1a
    def append_reads(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        self._ck(self.L.clb_append_reads(self.h, _np_ptr(bases), _np_ptr(offsets), n, 0))
        self.n_reads += n

    def append_quals(self, quals: np.ndarray):
        """Qualities of reads already appended (same order, no separators): they stay on the device and the quality encoders take
        quals=None.  Runs on the stage-3 stream, so a second host thread may call it while stages 1 / 2 run (clb_append_quals)."""
        q = np.ascontiguousarray(quals, np.uint8)
        self._ck(self.L.clb_append_quals(self.h, _np_ptr(q), len(q), 0))

    def append_reads_device(self, bases_ptr: int, offsets_ptr: int, n_reads: int):
        self._ck(self.L.clb_append_reads(self.h, C.c_void_p(bases_ptr), C.c_void_p(offsets_ptr), n_reads, 1))
        self.n_reads += n_reads

    def counts_export(self, part=0, n_parts=1):
        n = C.c_uint64()
        self._ck(self.L.clb_counts_size(self.h, part, n_parts, C.byref(n)))
        km = np.zeros(max(1, n.value), np.uint64)
        ct = np.zeros(max(1, n.value), np.uint32)
        m = C.c_uint64()
        self._ck(self.L.clb_counts_export(self.h, part, n_parts, _np_ptr(km), _np_ptr(ct), n.value, C.byref(m), 0))
        return km[:m.value], ct[:m.value]

    def counts_size(self, part, n_parts):
        n = C.c_uint64()
        self._ck(self.L.clb_counts_size(self.h, part, n_parts, C.byref(n)))
        return n.value

    def counts_export_device(self, part, n_parts, kmers_ptr, counts_ptr, cap):
        m = C.c_uint64()
        self._ck(self.L.clb_counts_export(self.h, part, n_parts, C.c_void_p(kmers_ptr), C.c_void_p(counts_ptr), cap, C.byref(m), 1))
        return m.value

    def counts_sizes(self, n_parts):
        """Entries of every partition of the count table, one pass (clb_counts_sizes)."""
        sizes = np.zeros(n_parts, np.uint64)
        self._ck(self.L.clb_counts_sizes(self.h, n_parts, _np_ptr(sizes)))
        return [int(x) for x in sizes]

    def counts_export_all_device(self, n_parts, first, kmers_ptr, counts_ptr, cap):
        """Every partition at kmers / counts [first[p] ..) (device buffers), one pass (clb_counts_export_all)."""
        f = np.ascontiguousarray(first, np.uint64)
        self._ck(self.L.clb_counts_export_all(self.h, n_parts, _np_ptr(f), C.c_void_p(kmers_ptr), C.c_void_p(counts_ptr), cap))

    def counts_merge_device(self, kmers_ptr, counts_ptr, n, n_reads_remote=0):
        self._ck(self.L.clb_counts_merge(self.h, C.c_void_p(kmers_ptr), C.c_void_p(counts_ptr), n, n_reads_remote, 1))

    def filter_size(self):
        n = C.c_uint64()
        st = self.L.clb_filter_list(self.h, None, None, 0, C.byref(n), 0)
        if st not in (0, 6):
            self._ck(st)
        return n.value

    def filter_list_device(self, kmers_ptr, counts_ptr, cap):
        n = C.c_uint64()
        self._ck(self.L.clb_filter_list(self.h, C.c_void_p(kmers_ptr), C.c_void_p(counts_ptr), cap, C.byref(n), 1))
        return n.value

    def filter_import_device(self, kmers_ptr, counts_ptr, n, stats=None):
        ks = KmerStats(*[stats[k] for k, _ in KmerStats._fields_]) if stats is not None else None
        self._ck(self.L.clb_filter_import(self.h, C.c_void_p(kmers_ptr), C.c_void_p(counts_ptr), n, C.byref(ks) if ks else None, 1))

    def counts_reset(self):
        self._ck(self.L.clb_counts_reset(self.h))

    def counts_merge(self, kmers, counts, n_reads_remote=0):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint32)
        self._ck(self.L.clb_counts_merge(self.h, _np_ptr(kmers), _np_ptr(counts), len(kmers), n_reads_remote, 0))

    def count_finalize(self):
        st = KmerStats()
        self._ck(self.L.clb_count_finalize(self.h, C.byref(st)))
        return st.as_dict()

    def filter_list(self):
        n = C.c_uint64()
        st = self.L.clb_filter_list(self.h, None, None, 0, C.byref(n), 0)
        if st not in (0, 6):
            self._ck(st)
        km = np.zeros(max(1, n.value), np.uint64)
        ct = np.zeros(max(1, n.value), np.uint32)
        self._ck(self.L.clb_filter_list(self.h, _np_ptr(km), _np_ptr(ct), n.value, C.byref(n), 0))
        return km[:n.value], ct[:n.value]

    def filter_import(self, kmers, counts, stats=None):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint32)
        ks = None
        if stats is not None:
            ks = KmerStats(*[stats[k] for k, _ in KmerStats._fields_])
        self._ck(self.L.clb_filter_import(self.h, _np_ptr(kmers), _np_ptr(counts), len(kmers), C.byref(ks) if ks else None, 0))

    def filter_check(self, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        poss = np.zeros(len(kmers), np.uint8)
        pres = np.zeros(len(kmers), np.uint8)
        self._ck(self.L.clb_filter_check(self.h, _np_ptr(kmers), len(kmers), _np_ptr(poss), _np_ptr(pres)))
        return poss, pres

    # ---- stage 1b
    def graph_build(self, is_reference=None, n_pseudo=0):
        if is_reference is None:
            ptr = None
        else:
            is_reference = np.ascontiguousarray(is_reference, np.uint8)
            assert len(is_reference) == self.n_reads - self.n_context
            ptr = _np_ptr(is_reference)
        self._ck(self.L.clb_graph_build(self.h, ptr, n_pseudo))

    def graph_accepted(self):
        tot = C.c_uint64()
        self._ck(self.L.clb_graph_accepted_size(self.h, C.byref(tot)))
        off = np.zeros(self.n_reads + 1, np.uint64)
        km = np.zeros(max(1, tot.value), np.uint64)
        self._ck(self.L.clb_graph_accepted(self.h, _np_ptr(off), _np_ptr(km), tot.value))
        return off, km[:tot.value]

    def graph_candidates(self):
        mc = self.params.max_candidates
        cand = np.zeros(max(1, self.n_reads * mc), np.uint32)
        cn = np.zeros(max(1, self.n_reads), np.uint32)
        self._ck(self.L.clb_graph_candidates(self.h, _np_ptr(cand), _np_ptr(cn)))
        return cand[:self.n_reads * mc].reshape(self.n_reads, mc), cn[:self.n_reads]

    def graph_common(self):
        mc = self.params.max_candidates
        tot = C.c_uint64()
        self._ck(self.L.clb_graph_common_size(self.h, C.byref(tot)))
        off = np.zeros(max(1, self.n_reads * mc), np.uint64)
        cn = np.zeros(max(1, self.n_reads * mc), np.uint32)
        km = np.zeros(max(1, tot.value), np.uint64)
        self._ck(self.L.clb_graph_common(self.h, _np_ptr(off), _np_ptr(cn), _np_ptr(km), tot.value))
        return off, cn, km[:tot.value]

    def packed_read(self, read_id):
        n = C.c_uint64()
        st = self.L.clb_get_packed_read(self.h, read_id, None, 0, C.byref(n))
        if st not in (0, 6):
            self._ck(st)
        out = np.zeros(n.value, np.uint8)
        self._ck(self.L.clb_get_packed_read(self.h, read_id, _np_ptr(out), n.value, C.byref(n)))
        return out

    # ---- stage 2
    def edit_scripts(self, cases):
        """cases: list of (kind, ref symbols, enc symbols, ref tail byte, enc tail byte) -> list of script bytes."""
        parts, ref_off, ref_len, enc_off, enc_len, kinds, o = [], [], [], [], [], [], 0
        for kind, ref, enc, rt, et in cases:
            ref_off.append(o); ref_len.append(len(ref)); parts += [np.asarray(ref, np.uint8), np.array([rt], np.uint8)]; o += len(ref) + 1
            enc_off.append(o); enc_len.append(len(enc)); parts += [np.asarray(enc, np.uint8), np.array([et], np.uint8)]; o += len(enc) + 1
            kinds.append(kind)
        seqs = np.concatenate(parts) if parts else np.zeros(1, np.uint8)
        n = len(cases)
        cap = int(sum(ref_len) + sum(enc_len) + 2 * n + 16)
        out = np.zeros(cap, np.uint8)
        out_off = np.zeros(n + 1, np.uint64)
        a = [np.array(x, t) for x, t in ((ref_off, np.uint64), (ref_len, np.uint32), (enc_off, np.uint64), (enc_len, np.uint32), (kinds, np.uint32))]
        self._ck(self.L.clb_edit_scripts(self.h, _np_ptr(seqs), len(seqs), *[_np_ptr(x) for x in a], n, _np_ptr(out_off), _np_ptr(out), cap))
        return [out[int(out_off[i]):int(out_off[i + 1])].tobytes() for i in range(n)]

    def encode(self, p, pack_sizes=None, keep_candidates=False):
        """Stage 2 over all appended reads.  p: dict with the keys of a golden params.txt (anchor_len, min_part_len_alt, ...)."""
        prm = p if isinstance(p, EncodeParams) else EncodeParams(
            p["anchor_len"], p["min_part_len_alt"], p["max_recurence"], p["min_anchors"],
            float(p["min_mmer_frac"]), float(p["min_mmer_force"]), float(p["max_matches_mult"]), float(p["es_cost_mult"]))
        if keep_candidates:
            self._ck(self.L.clb_encode_keep_candidates(self.h, 1))
        if pack_sizes is None:
            self._ck(self.L.clb_encode(self.h, C.byref(prm), None, 0))
        else:
            ps = np.ascontiguousarray(pack_sizes, np.uint32)
            self._ck(self.L.clb_encode(self.h, C.byref(prm), _np_ptr(ps), len(ps)))

    def encode_size(self):
        n = C.c_uint64()
        self._ck(self.L.clb_encode_size(self.h, C.byref(n)))
        return n.value

    def encoded(self, n_reads):
        """-> (es_off[n_reads + 1], CompactES bytes) on the host."""
        tot = self.encode_size()
        off = np.zeros(n_reads + 1, np.uint64)
        es = np.zeros(max(tot, 1), np.uint8)
        self._ck(self.L.clb_encode_get(self.h, _np_ptr(off), _np_ptr(es), tot, 0))
        return off, es[:tot]

    def encode_candidates(self, n_reads):
        """Parity tap -> per read a list of (ref_id, rev, tot, [(len, pos_enc, pos_ref), ...])."""
        n = C.c_uint64()
        self._ck(self.L.clb_encode_candidates_size(self.h, C.byref(n)))
        off = np.zeros(n_reads + 1, np.uint64)
        data = np.zeros(max(n.value, 1), np.uint32)
        self._ck(self.L.clb_encode_candidates(self.h, _np_ptr(off), _np_ptr(data), n.value))
        out = []
        for i in range(n_reads):
            rec, p, e = [], int(off[i]), int(off[i + 1])
            while p < e:
                ref_id, rev, tot, na = (int(x) for x in data[p:p + 4]); p += 4
                rec.append((ref_id, rev, tot, [tuple(int(x) for x in data[p + 3 * k:p + 3 * k + 3]) for k in range(na)]))
                p += 3 * na
            out.append(rec)
        return out

    # ---- multi-GPU: global reference-read set
    def append_context_reads(self, bases, offsets, n_reads=None, on_device=False):
        """Reference reads of earlier shards: ids in front of this context's own reads, not counted, never queried or encoded."""
        if on_device:
            self._ck(self.L.clb_append_context_reads(self.h, C.c_void_p(bases), C.c_void_p(offsets), n_reads, 1))
        else:
            b = np.ascontiguousarray(bases, np.uint8); o = np.ascontiguousarray(offsets, np.uint64)
            n_reads = len(o) - 1
            self._ck(self.L.clb_append_context_reads(self.h, _np_ptr(b) if len(b) else None, _np_ptr(o), n_reads, 0))
        self.n_reads += n_reads
        self.n_context += n_reads

    def reads_have_n(self, n_reads):
        out = np.zeros(max(n_reads, 1), np.uint8)
        self._ck(self.L.clb_reads_have_n(self.h, _np_ptr(out)))
        return out[:n_reads]

    def reads_export(self, read_ids, total_bases, device_ptr=None):
        """ASCII bases of the listed reads back to back -> numpy array (or written to device_ptr)."""
        ids = np.ascontiguousarray(read_ids, np.uint32)
        if device_ptr is not None:
            self._ck(self.L.clb_reads_export(self.h, _np_ptr(ids) if len(ids) else None, len(ids), C.c_void_p(device_ptr), int(total_bases), 1))
            return None
        out = np.zeros(max(int(total_bases), 1), np.uint8)
        self._ck(self.L.clb_reads_export(self.h, _np_ptr(ids) if len(ids) else None, len(ids), _np_ptr(out), int(total_bases), 0))
        return out[:int(total_bases)]

    # ---- stage 3
    def dna_encode(self, level, pack_sizes=None):
        """DNA / edit-script stream of all reads from the tuples of encode() (native container DB01)."""
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        self._ck(self.L.clb_dna_encode(self.h, level, None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))

    def dna_stream(self):
        """-> (container bytes, bytes of its table header)."""
        n, h = C.c_uint64(), C.c_uint64()
        self._ck(self.L.clb_dna_size(self.h, C.byref(n), C.byref(h)))
        out = np.zeros(max(n.value, 1), np.uint8)
        self._ck(self.L.clb_dna_get(self.h, _np_ptr(out), n.value, 0))
        return out[:n.value], h.value

    def qual_encode(self, n_bins, thresholds, level, quals, offsets, pack_sizes=None, on_device=False):
        """Quality stream of all appended reads (native container QB01).  quals/offsets: numpy arrays, or device pointers (ints)."""
        prm = QualParams()
        prm.n_bins, prm.level = n_bins, level
        for i, t in enumerate(thresholds):
            prm.thresholds[i] = t
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        if quals is None:      # resident (append_quals)
            qp, op = None, None
        elif on_device:
            qp, op = C.c_void_p(quals), C.c_void_p(offsets)
        else:
            q = np.ascontiguousarray(quals, np.uint8); o = np.ascontiguousarray(offsets, np.uint64)
            qp, op = _np_ptr(q), _np_ptr(o)
        self._ck(self.L.clb_qual_encode(self.h, C.byref(prm), qp, op, int(on_device), None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))

    def qual_encode_original(self, source, level, quals, offsets, pack_sizes=None, on_device=False):
        """Lossless quality stream (-q org; native container QO01).  source: 0 ONT, 1 CLR, 2 HiFi."""
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        if on_device:
            qp, op = C.c_void_p(quals), C.c_void_p(offsets)
        else:
            q = np.ascontiguousarray(quals, np.uint8); o = np.ascontiguousarray(offsets, np.uint64)
            qp, op = _np_ptr(q), _np_ptr(o)
        self._ck(self.L.clb_qual_encode_original(self.h, source, level, qp, op, int(on_device), None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))

    def qual_size(self):
        n = C.c_uint64()
        self._ck(self.L.clb_qual_size(self.h, C.byref(n)))
        return n.value

    def qual_stream(self):
        n = self.qual_size()
        out = np.zeros(max(n, 1), np.uint8)
        self._ck(self.L.clb_qual_get(self.h, _np_ptr(out), n, 0))
        return out[:n]

    def hdr_encode(self, headers=None, plus_id=None, pack_sizes=None, *, bytes_=None, offsets=None, n=None, on_device=False):
        """Header stream (native container HB01).  headers: list of bytes objects — or bytes_/offsets (numpy arrays, or device
        pointers with n and on_device=True).  plus_id: per-header flag "the '+' line repeats the header" (None = never)."""
        if headers is not None:
            bytes_ = np.frombuffer(b"".join(headers), np.uint8)
            offsets = np.zeros(len(headers) + 1, np.uint64)
            offsets[1:] = np.cumsum([len(h) for h in headers], dtype=np.uint64)
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        if on_device:
            bp, op, pp = C.c_void_p(bytes_), C.c_void_p(offsets), (None if plus_id is None else C.c_void_p(plus_id))
        else:
            b = np.ascontiguousarray(bytes_, np.uint8); o = np.ascontiguousarray(offsets, np.uint64); n = len(o) - 1
            pl = None if plus_id is None else np.ascontiguousarray(plus_id, np.uint8)
            bp, op, pp = (_np_ptr(b) if len(b) else None), _np_ptr(o), (None if pl is None else _np_ptr(pl))
        self._ck(self.L.clb_hdr_encode(self.h, bp, op, pp, n, int(on_device), None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))

    def hdr_stream(self):
        """-> (container bytes, bytes of its table part)."""
        n, h = C.c_uint64(), C.c_uint64()
        self._ck(self.L.clb_hdr_size(self.h, C.byref(n), C.byref(h)))
        out = np.zeros(max(n.value, 1), np.uint8)
        self._ck(self.L.clb_hdr_get(self.h, _np_ptr(out), n.value, 0))
        return out[:n.value], h.value

    # ---- compat streams: the reference's own parts (stage3_exact.cu) ----
    QMODES = {"org": 0, "5-avg": 1, "4-avg": 2, "2-avg": 3, "5-fix": 4, "4-fix": 5, "2-fix": 6, "avg": 7, "none": 8}

    def xdna_encode(self, level, pack_sizes=None):
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        self._ck(self.L.clb_xdna_encode(self.h, level, None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))
        return self.xstream(0)

    def xqual_encode(self, mode, source, level, thresholds, quals, offsets, pack_sizes=None):
        """mode: a -q name of the reference; source 0 ONT / 1 CLR / 2 HiFi; thresholds: the forward thresholds of the binned modes."""
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        t = np.zeros(8, np.uint32); t[:len(thresholds)] = thresholds
        q = np.ascontiguousarray(quals, np.uint8); o = np.ascontiguousarray(offsets, np.uint64)
        self._ck(self.L.clb_xqual_encode(self.h, self.QMODES[mode], source, level, _np_ptr(t), _np_ptr(q), _np_ptr(o), 0, None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))
        return self.xstream(1)

    def xhdr_encode(self, headers, plus_id=None, pack_sizes=None):
        b = np.frombuffer(b"".join(headers), np.uint8)
        o = np.zeros(len(headers) + 1, np.uint64)
        o[1:] = np.cumsum([len(h) for h in headers], dtype=np.uint64)
        ps = None if pack_sizes is None else np.ascontiguousarray(pack_sizes, np.uint32)
        pl = None if plus_id is None else np.ascontiguousarray(plus_id, np.uint8)
        self._ck(self.L.clb_xhdr_encode(self.h, _np_ptr(b) if len(b) else None, _np_ptr(o), None if pl is None else _np_ptr(pl), len(headers), 0,
                                        None if ps is None else _np_ptr(ps), 0 if ps is None else len(ps)))
        return self.xstream(2)

    def xstream(self, which):
        """-> the parts of compat stream `which` (0 dna, 1 qual, 2 header) as a list of bytes."""
        n, k = C.c_uint64(), C.c_uint32()
        self._ck(self.L.clb_xstream_size(self.h, which, C.byref(n), C.byref(k)))
        out = np.zeros(max(n.value, 1), np.uint8); sizes = np.zeros(max(k.value, 1), np.uint64)
        self._ck(self.L.clb_xstream_get(self.h, which, _np_ptr(out), n.value, _np_ptr(sizes), 0))
        parts, at = [], 0
        for i in range(k.value):
            parts.append(out[at:at + int(sizes[i])].tobytes()); at += int(sizes[i])
        return parts

    def dna_stream_size(self):
        n, h = C.c_uint64(), C.c_uint64()
        self._ck(self.L.clb_dna_size(self.h, C.byref(n), C.byref(h)))
        return n.value

    def hdr_stream_size(self):
        n, h = C.c_uint64(), C.c_uint64()
        self._ck(self.L.clb_hdr_size(self.h, C.byref(n), C.byref(h)))
        return n.value

    def stream_into(self, which, host_ptr, cap):
        """Copy a finished stream ("dna" / "qual" / "hdr") into caller-owned HOST memory (e.g. pinned); -> its size in bytes."""
        n, h = C.c_uint64(), C.c_uint64()
        if which == "qual":
            self._ck(self.L.clb_qual_size(self.h, C.byref(n)))
        else:
            self._ck(getattr(self.L, f"clb_{which}_size")(self.h, C.byref(n), C.byref(h)))
        self._ck(getattr(self.L, f"clb_{which}_get")(self.h, C.c_void_p(host_ptr), cap, 0))
        return n.value

    def profile_enable(self, on=True):
        self._ck(self.L.clb_profile_enable(self.h, int(on)))

    def profile(self):
        """-> {kernel class: (total ms, launches)} measured with CUDA events on the context's stream."""
        out = {}
        for k in KERNEL_CLASSES:
            ms, n = C.c_double(), C.c_uint64()
            self._ck(self.L.clb_profile_get(self.h, k.encode(), C.byref(ms), C.byref(n)))
            out[k] = (ms.value, n.value)
        return out

    @property
    def kernel_launches(self):
        return int(self.L.clb_kernel_launches(self.h))
