"""Deterministic synthetic long-read FASTQ generators (SURVEY.md §8d / BASELINE.md §2 recipe).

One ``numpy.random.default_rng(seed)`` is consumed in a fixed order so the same (seed, sizes) always
gives the same bytes: genome, then per read length, start, strand, error classes, qualities, header.
The ONT profile is the BASELINE.md recipe verbatim; HiFi / CLR profiles change error rate, error mix,
quality distribution and header style as SURVEY.md §8(d) describes.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

PROFILES = {
    # err, (sub, del, ins) split of err, quality (mean, sd, lo, hi)
    "ont": dict(err=0.10, split=(0.4, 0.3, 0.3), q=(12, 5, 1, 40)),
    "clr": dict(err=0.13, split=(2 / 13, 4 / 13, 7 / 13), q=(10, 4, 1, 30)),
    "hifi": dict(err=0.002, split=(0.2, 0.4, 0.4), q=(80, 15, 1, 93)),
}


class SynthReads:
    """Container: concatenated ASCII bases / qualities + offsets + headers (all numpy)."""

    def __init__(self, bases, quals, offsets, headers):
        self.bases = bases          # uint8 ASCII 'ACGT(N)', concatenated, no separators
        self.quals = quals          # uint8 ASCII phred+33, same layout
        self.offsets = offsets      # uint64[n+1]
        self.headers = headers      # list[bytes] without leading '@'

    @property
    def n_reads(self):
        return len(self.offsets) - 1

    @property
    def n_bases(self):
        return int(self.offsets[-1])

    def fastq_bytes(self) -> int:
        return sum(len(h) + 1 for h in self.headers) + 2 * self.n_bases + 5 * self.n_reads

    def write_fastq(self, path):
        with open(path, "wb") as f:
            off = self.offsets
            for i, h in enumerate(self.headers):
                s, e = int(off[i]), int(off[i + 1])
                f.write(b"@" + h + b"\n")
                f.write(self.bases[s:e].tobytes())
                f.write(b"\n+\n")
                f.write(self.quals[s:e].tobytes())
                f.write(b"\n")


def generate(n_reads: int, genome_len: int, mean_len: int, seed: int, profile: str = "ont",
             n_frac: float = 0.0, min_len: int = 200) -> SynthReads:
    p = PROFILES[profile]
    e = p["err"]
    t_sub = p["split"][0] * e
    t_del = t_sub + p["split"][1] * e
    qm, qs, qlo, qhi = p["q"]
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    seqs, quals, headers = [], [], []
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    for i in range(n_reads):
        ln = int(np.clip(rng.gamma(2.0, mean_len / 2.0), min_len, genome_len - 1))
        start = int(rng.integers(0, genome_len - ln))
        frag = genome[start:start + ln]
        if rng.random() < 0.5:
            frag = 3 - frag[::-1]
        r = rng.random(ln)
        sub = r < t_sub
        dele = (r >= t_sub) & (r < t_del)
        ins = (r >= t_del) & (r < e)
        frag = frag.copy()
        nsub = int(sub.sum())
        if nsub:
            frag[sub] = (frag[sub] + rng.integers(1, 4, nsub, dtype=np.uint8)) & 3
        rep = np.ones(ln, dtype=np.int64)
        rep[dele] = 0
        rep[ins] = 2
        out = np.repeat(frag, rep)
        nins = int(ins.sum())
        if nins:
            # the second copy of every duplicated position is overwritten by a random base
            ends = np.cumsum(rep)[ins] - 1
            out[ends] = rng.integers(0, 4, nins, dtype=np.uint8)
        q = np.clip(np.rint(rng.normal(qm, qs, len(out))), qlo, qhi).astype(np.uint8) + 33
        asc = _ACGT[out]
        if n_frac > 0 and rng.random() < n_frac and len(asc) > 10:
            npos = rng.integers(0, len(asc), max(1, len(asc) // 2000))
            asc = asc.copy()
            asc[npos] = ord("N")
        if profile == "ont":
            ch = int(rng.integers(1, 513))
            h = b"read_%d ch=%d start_time=2020-01-01T00:%02d:%02dZ" % (i, ch, (i // 60) % 60, i % 60)
        elif profile == "hifi":
            h = b"m64011_190830_220126/%d/ccs" % (i * 3 + 17)
        else:
            h = b"m54238_180901_011437/%d/0_%d" % (i * 5 + 11, len(out))
        seqs.append(asc)
        quals.append(q)
        headers.append(h)
        offsets[i + 1] = offsets[i] + np.uint64(len(out))
    return SynthReads(np.concatenate(seqs), np.concatenate(quals), offsets, headers)


# ----------------------------------------------------------------------------------------------------------------------
# FASTQ files of the BASELINE configurations at file scale (tools/ratio_check.py, bench.py): same error / quality / header
# model as generate(), read lengths per SURVEY.md §8d (ONT / CLR: gamma(2); HiFi: normal), written by a pool of processes.
# Deterministic in (profile, n_reads, genome_len, mean_len, seed) — not in the number of workers (chunks are fixed-size).
# ----------------------------------------------------------------------------------------------------------------------
_FILE_CHUNK = 2000
_G = {}


def _file_chunk(args):
    profile, c, n_reads, mean_len, seed = args
    genome = _G["genome"]
    G = len(genome)
    p = PROFILES[profile]
    e = p["err"]
    t_sub = p["split"][0] * e
    t_del = t_sub + p["split"][1] * e
    qm, qs, qlo, qhi = p["q"]
    rng = np.random.default_rng([seed, c])
    lo, hi = c * _FILE_CHUNK, min(n_reads, (c + 1) * _FILE_CHUNK)
    n = hi - lo
    if profile == "hifi":
        lens = np.clip(rng.normal(mean_len, mean_len / 7.5, n), 1000, None).astype(np.int64)
    else:
        lens = np.clip(rng.gamma(2.0, mean_len / 2.0, n), 200, 200000).astype(np.int64)
    lens = np.minimum(lens, G - 1)
    start = (rng.random(n) * (G - lens)).astype(np.int64)
    rev = rng.random(n) < 0.5
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(lens)
    rid = np.repeat(np.arange(n), lens)
    within = np.arange(off[-1]) - off[rid]
    src = np.where(rev[rid], start[rid] + lens[rid] - 1 - within, start[rid] + within)
    frag = genome[src]
    frag = np.where(rev[rid], 3 - frag, frag).astype(np.uint8)
    r = rng.random(len(frag))
    sub = r < t_sub
    rep = np.ones(len(frag), np.int64)
    rep[(r >= t_sub) & (r < t_del)] = 0
    ins = (r >= t_del) & (r < e)
    rep[ins] = 2
    frag[sub] = (frag[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
    out = np.repeat(frag, rep)
    out[np.cumsum(rep)[ins] - 1] = rng.integers(0, 4, int(ins.sum()), dtype=np.uint8)
    out_len = np.add.reduceat(rep, off[:-1]) if n else np.zeros(0, np.int64)
    out_len = np.maximum(out_len, 0)
    if profile == "hifi":      # mostly '~' (93) with dips
        q = np.full(len(out), 93, np.int64)
        dip = rng.random(len(out)) < 0.08
        q[dip] = np.clip(np.rint(rng.normal(40, 20, int(dip.sum()))), 1, 93).astype(np.int64)
    else:
        q = np.clip(np.rint(rng.normal(qm, qs, len(out))), qlo, qhi).astype(np.int64)
    q = (q + 33).astype(np.uint8)
    asc = _ACGT[out]
    ooff = np.zeros(n + 1, np.int64)
    ooff[1:] = np.cumsum(out_len)
    chs = rng.integers(1, 513, n)
    parts = []
    n_bases = 0
    for j in range(n):
        i = lo + j
        a, b = int(ooff[j]), int(ooff[j + 1])
        if b == a:      # every base deleted: keep the record non-empty
            continue
        if profile == "ont":
            h = b"@read_%d ch=%d start_time=2020-01-01T%02d:%02d:%02dZ\n" % (i, chs[j], (i // 3600) % 24, (i // 60) % 60, i % 60)
        elif profile == "hifi":
            h = b"@m64011_190830_220126/%d/ccs\n" % (i * 3 + 17)
        else:
            h = b"@m54238_180901_011437/%d/0_%d\n" % (i * 5 + 11, b - a)
        parts += [h, asc[a:b].tobytes(), b"\n+\n", q[a:b].tobytes(), b"\n"]
        n_bases += b - a
    return b"".join(parts), n_bases


def generate_file(path: str, profile: str, n_reads: int, genome_len: int, mean_len: int, seed: int, workers: int | None = None, fasta_genome: str | None = None):
    """Writes a FASTQ of `n_reads` synthetic reads; returns (file bytes, bases).  fasta_genome: also write the genome (for -G)."""
    import multiprocessing as mp
    import os
    _G["genome"] = np.random.default_rng([seed, 1 << 30]).integers(0, 4, genome_len, dtype=np.uint8)
    if fasta_genome:
        with open(fasta_genome, "wb") as f:
            f.write(b">synthetic_genome\n")
            asc = _ACGT[_G["genome"]]
            for a in range(0, genome_len, 80 * 100000):
                blk = asc[a:a + 80 * 100000]
                rows = [blk[r:r + 80].tobytes() for r in range(0, len(blk), 80)]
                f.write(b"\n".join(rows) + b"\n")
    n_chunks = (n_reads + _FILE_CHUNK - 1) // _FILE_CHUNK
    tasks = [(profile, c, n_reads, mean_len, seed) for c in range(n_chunks)]
    workers = workers or min(32, os.cpu_count() or 1)
    total = bases = 0
    with open(path, "wb") as f:
        if workers <= 1 or n_chunks <= 1:
            it = map(_file_chunk, tasks)
            for blob, nb in it:
                f.write(blob); total += len(blob); bases += nb
        else:
            with mp.get_context("fork").Pool(workers) as pool:      # fork: the genome is shared copy-on-write
                for blob, nb in pool.imap(_file_chunk, tasks, chunksize=1):
                    f.write(blob); total += len(blob); bases += nb
    _G.clear()
    return total, bases
