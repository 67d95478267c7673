#!/usr/bin/env python
"""Generate tests/golden/edit_scripts.bin.gz: alignment + canonicalisation cases answered by the
reference itself (oracle/_ref/ref_edit_script = the reference's edlib + edit_script.h).  Build container only."""
import gzip
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def mutate(rng, s, err):
    out = []
    for b in s:
        r = rng.random()
        if r < 0.4 * err:
            out.append((b + rng.integers(1, 4)) & 3)
        elif r < 0.7 * err:
            continue
        elif r < err:
            out.append(b)
            out.append(rng.integers(0, 4))
        else:
            out.append(b)
    return np.array(out, np.uint8)


def cases():
    rng = np.random.default_rng(2024)
    out = []
    sizes = [0, 1, 2, 3, 5, 8, 13, 14, 15, 16, 20, 31, 33, 63, 64, 65, 100, 127, 128, 129, 200, 300, 500, 700, 1000, 1500]
    for kind in (0, 1, 2):
        for n in sizes:
            for err in (0.0, 0.05, 0.15, 0.4):
                ref = rng.integers(0, 4, n).astype(np.uint8)
                if rng.random() < 0.3 and n > 4:          # homopolymer-rich: exercises refactor_edit_script
                    ref = np.repeat(rng.integers(0, 4, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))[:n].astype(np.uint8)
                enc = mutate(rng, ref, err)
                if kind != 2 and n > 0:                   # flanks: the reference side is longer than what the read covers
                    extra = rng.integers(0, 4, int(rng.integers(0, 3 * n + 5))).astype(np.uint8)
                    ref = np.concatenate([extra, ref]) if kind == 0 else np.concatenate([ref, extra])
                out.append((kind, ref, enc))
    # unequal lengths, empty sides, unrelated sequences
    for kind in (0, 1, 2):
        for (a, b) in [(0, 7), (7, 0), (1, 50), (50, 1), (2, 2), (300, 20), (20, 300), (64, 640), (640, 64)]:
            out.append((kind, rng.integers(0, 4, a).astype(np.uint8), rng.integers(0, 4, b).astype(np.uint8)))
    # big ones: edlib switches from the stored-column traceback to Hirschberg (edlib.cpp:1191-1214)
    for kind, n, err in [(2, 2500, 0.1), (2, 4000, 0.12), (2, 6000, 0.05), (1, 3000, 0.1), (0, 3000, 0.1), (2, 9000, 0.15), (2, 1800, 0.3)]:
        ref = rng.integers(0, 4, n).astype(np.uint8)
        enc = mutate(rng, ref, err)
        if kind == 0:
            ref = np.concatenate([rng.integers(0, 4, 500).astype(np.uint8), ref])
        if kind == 1:
            ref = np.concatenate([ref, rng.integers(0, 4, 500).astype(np.uint8)])
        out.append((kind, ref, enc))
    # short flanks against unrelated reference symbols: edlib's SHW also scores the EMPTY target prefix (position -1 of its
    # wildcard-padded query) and lists it first, so a part no prefix aligns to better than |enc| insertions comes out as insertions only
    for kind in (0, 1):
        for el in (2, 3, 4, 5, 6, 8, 11):
            for rl in (2, 3, 4, 7, 16, 40):
                for _ in range(3):
                    out.append((kind, rng.integers(0, 4, rl).astype(np.uint8), rng.integers(0, 4, el).astype(np.uint8)))
        for el in (2, 3, 5):              # no symbol of the part occurs in the reference window at all
            out.append((kind, np.full(9, 1, np.uint8), np.full(el, 2, np.uint8)))
            out.append((kind, np.array([0, 1] * 6, np.uint8), np.array([2, 3] * el, np.uint8)[:el]))
    return out, rng


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_edit_script")
    if not os.path.exists(exe):
        sys.exit("build the oracle first: make -C oracle ref drivers")
    cs, rng = cases()
    blob = bytearray()
    tails = []
    for kind, ref, enc in cs:
        rt, et = int(rng.integers(0, 4)), int(rng.integers(0, 4))      # the byte that follows each part in its read
        if rng.random() < 0.2:
            rt = 255
        if rng.random() < 0.2:
            et = 255
        tails.append((rt, et))
        blob += struct.pack("<III", kind, len(ref), len(enc)) + ref.tobytes() + bytes([rt]) + enc.tobytes() + bytes([et])
    res = subprocess.run([exe], input=bytes(blob), stdout=subprocess.PIPE, check=True).stdout
    out = bytearray(struct.pack("<I", len(cs)))
    pos = 0
    for (kind, ref, enc), (rt, et) in zip(cs, tails):
        (n,) = struct.unpack_from("<I", res, pos)
        script = res[pos + 4:pos + 4 + n]
        pos += 4 + n
        out += struct.pack("<IIIBBI", kind, len(ref), len(enc), rt, et, n) + ref.tobytes() + enc.tobytes() + script
    assert pos == len(res)
    path = os.path.join(ROOT, "tests", "golden", "edit_scripts.bin.gz")
    with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as f:
        f.write(bytes(out))
    print(len(cs), "cases,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
