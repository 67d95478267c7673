#!/usr/bin/env python
"""Fixtures for tests/test_host_compat.py: small archives written by the UNMODIFIED reference CLI (oracle/_ref/colord), one per
stream flavour its decoders have (compression level 1 / 2 / 3, every -q mode, FASTA, -i none, all-reference mode), with the SHA-1 of
what the reference's own `decompress` prints for each.  Build container only (needs oracle/_ref/colord); inputs are synthetic
(colord_b200.synth.generate), so nothing of the reference's data is stored.
Usage: python tests/golden/make_ref_archive_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from colord_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_archives")
ONT = dict(n_reads=120, genome_len=12000, mean_len=1500, seed=31, profile="ont", n_frac=0.05)
CLR = dict(n_reads=100, genome_len=12000, mean_len=1500, seed=32, profile="clr", n_frac=0.03)
HIFI = dict(n_reads=60, genome_len=30000, mean_len=3000, seed=33, profile="hifi", n_frac=0.0)
CASES = {
    "ont_default": (ONT, ["compress-ont"]),
    "ont_balanced_5avg": (ONT, ["compress-ont", "-p", "balanced", "-q", "5-avg"]),
    "ont_ratio_2avg": (ONT, ["compress-ont", "-p", "ratio", "-q", "2-avg"]),
    "clr_ratio_org": (CLR, ["compress-pbraw", "-p", "ratio", "-q", "org"]),
    "clr_default": (CLR, ["compress-pbraw"]),
    "hifi_default": (HIFI, ["compress-pbhifi"]),
    "hifi_balanced_org": (HIFI, ["compress-pbhifi", "-p", "balanced", "-q", "org"]),
    "ont_2fix": (ONT, ["compress-ont", "-q", "2-fix"]),
    "ont_4fix_thr_bal": (ONT, ["compress-ont", "-q", "4-fix", "-T", "5", "12", "20", "-p", "balanced"]),
    "ont_5fix": (ONT, ["compress-ont", "-q", "5-fix"]),
    "ont_avg": (ONT, ["compress-ont", "-q", "avg"]),
    "ont_none": (ONT, ["compress-ont", "-q", "none"]),
    "ont_id_none": (ONT, ["compress-ont", "-i", "none"]),
    "ont_fasta": (ONT, ["compress-ont"]),
    # reference-genome mode: the genome (GENOME below, three sequences with lower case, N runs and blank lines) stored in the archive / given again
    "ont_genome_stored": (ONT, ["compress-ont", "-G", "genome.fa", "-s"]),
    "ont_genome_checksum_bal": (ONT, ["compress-ont", "-G", "genome.fa", "-p", "balanced"]),
}


def write_genome(path, gen):
    """the genome the synthetic reads of `gen` were drawn from (synth.generate's first draw), cut into three FASTA records"""
    import numpy as np
    g = np.random.default_rng(gen["seed"]).integers(0, 4, gen["genome_len"], dtype=np.uint8)
    asc = np.frombuffer(b"ACGT", np.uint8)[g].tobytes().decode()
    a, b = len(asc) // 3, 2 * len(asc) // 3
    recs = [("chr1 first", asc[:a]), ("chr2", asc[a:b].lower()), ("chr3 with N", asc[b:b + 500] + "N" * 37 + asc[b + 500:])]
    with open(path, "w") as f:
        for name, seq in recs:
            f.write(">" + name + "\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + "\n")
            f.write("\n")


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "colord")
    os.makedirs(OUT, exist_ok=True)
    exp = {}
    for name, (gen, cli) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            s = synth.generate(**gen)
            inp = os.path.join(tmp, "in.fasta" if name == "ont_fasta" else "in.fastq")
            if name == "ont_fasta":
                with open(inp, "wb") as f:
                    for i, h in enumerate(s.headers):
                        f.write(b">" + h + b"\n" + s.bases[int(s.offsets[i]):int(s.offsets[i + 1])].tobytes() + b"\n")
            else:
                s.write_fastq(inp)
            if "-G" in cli:
                write_genome(os.path.join(tmp, "genome.fa"), gen)
            arc = os.path.join(OUT, name + ".colord")
            subprocess.run([exe, *cli, "-t", "2", inp, arc], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
            out = os.path.join(tmp, "out")
            subprocess.run([exe, "decompress", *(["-G", "genome.fa"] if "-G" in cli and "-s" not in cli else []), arc, out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
            data = open(out, "rb").read()
            exp[name] = dict(generator=gen, cli=cli, output_sha1=hashlib.sha1(data).hexdigest(), output_bytes=len(data),
                             lossless=("org" in cli and "-i" not in cli) or name == "ont_fasta", input_sha1=hashlib.sha1(open(inp, "rb").read()).hexdigest())
            print(name, os.path.getsize(arc), len(data))
    with open(os.path.join(OUT, "expected.json"), "w") as f:
        json.dump(exp, f, indent=1)


if __name__ == "__main__":
    main()
