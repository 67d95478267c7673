"""Reader of the reference's archive container (archive.cpp:100-197; layout in SURVEY.md appendix A).  Tests only."""
from __future__ import annotations


def _varint(buf, p):
    n = buf[p]
    return int.from_bytes(buf[p + 1:p + 1 + n], "big"), p + 1 + n


def read_parts(path):
    """-> {stream name: [(metadata, payload bytes), ...]} in part order."""
    with open(path, "rb") as f:
        raw = f.read()
    fl = int.from_bytes(raw[-8:], "little")
    foot = raw[len(raw) - 8 - fl:len(raw) - 8]
    n_streams, p = _varint(foot, 0)
    out = {}
    for _ in range(n_streams):
        z = foot.index(b"\0", p)
        name = foot[p:z].decode()
        p = z + 1
        n_parts, p = _varint(foot, p)
        _raw_size, p = _varint(foot, p)
        parts = []
        for _ in range(n_parts):
            off, p = _varint(foot, p)
            size, p = _varint(foot, p)
            md, q = _varint(raw, off)
            parts.append((md, raw[q:q + size]))
        out[name] = parts
    return out
