"""Load a golden case: regenerate its synthetic input and parse the reference's dumps (tests only)."""
from __future__ import annotations

import hashlib
import json
import os
from types import SimpleNamespace

import numpy as np

import refdump
from colord_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    d = os.path.join(GOLDEN, name)
    with open(os.path.join(d, "input.json")) as f:
        meta = json.load(f)
    s = synth.generate(**meta["generator"])
    # the generator must reproduce the exact input the reference saw (numpy Generator streams are not
    # guaranteed across numpy versions; fail loudly instead of comparing against the wrong reads)
    assert hashlib.sha1(s.bases.tobytes()).hexdigest() == meta["bases_sha1"], "synthetic input drifted"
    assert hashlib.sha1(s.offsets.tobytes()).hexdigest() == meta["offsets_sha1"], "synthetic input drifted"
    params = refdump.load_params(d)
    kmers, counts = refdump.load_kmers(d)
    reads, packs = refdump.load_reads(d)
    es, es_packs = refdump.load_es(d)
    has_n = np.array([r["has_n"] for r in reads], np.uint8)
    is_ref = np.array([r["is_ref"] for r in reads], np.uint8)
    return SimpleNamespace(name=name, meta=meta, reads_in=s, params=params, kmers=kmers, counts=counts,
                           reads=reads, packs=packs, es=es, es_packs=es_packs, has_n=has_n, is_ref=is_ref)


def load_edit_scripts():
    """-> list of (kind, ref symbols, enc symbols, ref tail byte, enc tail byte, reference's edit script bytes)."""
    import gzip
    import struct
    with gzip.open(os.path.join(GOLDEN, "edit_scripts.bin.gz"), "rb") as f:
        raw = f.read()
    (n,) = struct.unpack_from("<I", raw, 0)
    pos, out = 4, []
    for _ in range(n):
        kind, rl, el, rt, et, sn = struct.unpack_from("<IIIBBI", raw, pos)
        pos += 18
        ref = np.frombuffer(raw, np.uint8, rl, pos).copy(); pos += rl
        enc = np.frombuffer(raw, np.uint8, el, pos).copy(); pos += el
        out.append((kind, ref, enc, rt, et, raw[pos:pos + sn])); pos += sn
    return out


def load_qual_golden(name):
    """-> (bases u8 ASCII, quals u8, expected lossy quals u8, offsets u64) from tests/golden/qual_<name>.bin.gz."""
    import gzip
    import struct
    with gzip.open(os.path.join(GOLDEN, f"qual_{name}.bin.gz"), "rb") as f:
        raw = f.read()
    (n,) = struct.unpack_from("<I", raw, 0)
    lens = np.frombuffer(raw, np.uint32, n, 4)
    tot = int(lens.sum())
    p = 4 + 4 * n
    bases = np.frombuffer(raw, np.uint8, tot, p).copy()
    quals = np.frombuffer(raw, np.uint8, tot, p + tot).copy()
    quan = np.frombuffer(raw, np.uint8, tot, p + 2 * tot).copy()
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    return bases, quals, quan, off


def load_hdr_golden():
    """-> {case: (list of header bytes, header-stream bytes of the unmodified reference)} from tests/golden/headers.json.gz
    (written by tests/golden/make_hdr_golden.py; the synthetic case is regenerated from its recipe)."""
    import gzip
    import importlib.util
    with gzip.open(os.path.join(GOLDEN, "headers.json.gz"), "rt") as f:
        raw = json.load(f)
    spec = importlib.util.spec_from_file_location("make_hdr_golden", os.path.join(GOLDEN, "make_hdr_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name, rec in raw.items():
        hs = mod.synth_headers() if rec["headers"] is None else [h.encode("latin-1") for h in rec["headers"]]
        out[name] = (hs, rec["ref_header_stream_bytes"])
    return out
