"""Parser for the per-stage dumps written by oracle/_ref/ref_stage_dump (see oracle/ref_stage_dump.cpp)."""
from __future__ import annotations

import gzip
import os
import struct

import numpy as np


def _open(path):
    if os.path.exists(path):
        return open(path, "rb")
    return gzip.open(path + ".gz", "rb")


def load_params(d):
    out = {}
    with _open(os.path.join(d, "params.txt")) as f:
        for line in f.read().decode().split():
            k, v = line.split("=")
            out[k] = float(v) if "." in v or "e" in v else int(v)
    return out


def load_kmers(d):
    with _open(os.path.join(d, "kmers.bin")) as f:
        raw = f.read()
    (n,) = struct.unpack_from("<Q", raw, 0)
    rec = np.frombuffer(raw, dtype=np.dtype([("kmer", "<u8"), ("count", "<u4")]), offset=8, count=n)
    return rec["kmer"].copy(), rec["count"].copy()


def load_reads(d):
    """-> list of dict(id, has_n, is_ref, length, acc[u64], cands[u32], common[list[u64 array]]) and pack sizes."""
    with _open(os.path.join(d, "reads.bin")) as f:
        raw = f.read()
    pos, reads, packs = 0, [], []
    while pos < len(raw):
        marker, pack_id, n = struct.unpack_from("<III", raw, pos)
        assert marker == 0xFFFFFFFF
        pos += 12
        packs.append(n)
        for _ in range(n):
            rid, has_n, is_ref, length, n_acc = struct.unpack_from("<IBBII", raw, pos)
            pos += 14
            acc = np.frombuffer(raw, dtype="<u8", count=n_acc, offset=pos).copy()
            pos += 8 * n_acc
            (n_c,) = struct.unpack_from("<I", raw, pos)
            pos += 4
            cands = np.frombuffer(raw, dtype="<u4", count=n_c, offset=pos).copy()
            pos += 4 * n_c
            (n_cm,) = struct.unpack_from("<I", raw, pos)
            pos += 4
            common = []
            for _ in range(n_cm):
                (m,) = struct.unpack_from("<I", raw, pos)
                pos += 4
                common.append(np.frombuffer(raw, dtype="<u8", count=m, offset=pos).copy())
                pos += 8 * m
            reads.append(dict(id=rid, has_n=has_n, is_ref=is_ref, length=length, acc=acc, cands=cands, common=common))
    return reads, packs


def load_es(d):
    """-> list of bytes (CompactES byte strings, one per read, input order) and pack sizes."""
    with _open(os.path.join(d, "es.bin")) as f:
        raw = f.read()
    pos, out, packs = 0, [], []
    while pos < len(raw):
        marker, n = struct.unpack_from("<II", raw, pos)
        assert marker == 0xFFFFFFFF
        pos += 8
        packs.append(n)
        for _ in range(n):
            (m,) = struct.unpack_from("<I", raw, pos)
            pos += 4
            out.append(raw[pos:pos + m])
            pos += m
    return out, packs
