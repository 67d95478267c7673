#!/usr/bin/env python
"""Golden vectors for the entropy-coded streams: the archive parts the UNMODIFIED reference CLI (oracle/_ref/colord) writes.

Build container only.  For every case the deterministic synthetic FASTQ is regenerated (colord_b200.synth), the stock binary
compresses it, and per stream ("dna", "qual", "header") the list of parts is recorded as (metadata, size, SHA-1 of the payload) in
tests/golden/streams.json.  Cases named after a directory of tests/golden use that case's input and CLI (their CompactES dumps are
the DNA coder's input); the `q_*` cases vary the quality mode on the ont_mem input; `multi_*` are inputs of several packs.
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from colord_b200 import synth  # noqa: E402
import colord_archive  # noqa: E402

ONT = dict(n_reads=400, genome_len=30000, mean_len=3000, seed=11, profile="ont", n_frac=0.05)
CASES = {
    "ont_mem": (ONT, ["compress-ont"]),
    "ont_bal": (ONT, ["compress-ont", "-p", "balanced"]),
    "clr_ratio": (dict(n_reads=300, genome_len=30000, mean_len=3000, seed=12, profile="clr", n_frac=0.03), ["compress-pbraw", "-p", "ratio"]),
    "hifi": (dict(n_reads=200, genome_len=100000, mean_len=6000, seed=13, profile="hifi", n_frac=0.0), ["compress-pbhifi"]),
    "q_org": (ONT, ["compress-ont", "-q", "org"]),
    "q_org_bal": (ONT, ["compress-ont", "-q", "org", "-p", "balanced"]),
    "q_org_ratio": (ONT, ["compress-ont", "-q", "org", "-p", "ratio"]),
    "q_2avg": (ONT, ["compress-ont", "-q", "2-avg"]),
    "q_5avg": (ONT, ["compress-ont", "-q", "5-avg"]),
    "q_2fix": (ONT, ["compress-ont", "-q", "2-fix"]),
    "q_4fix": (ONT, ["compress-ont", "-q", "4-fix"]),
    "q_5fix": (ONT, ["compress-ont", "-q", "5-fix"]),
    "q_avg": (ONT, ["compress-ont", "-q", "avg"]),
    "q_none": (ONT, ["compress-ont", "-q", "none"]),
    "q_4fix_thr": (ONT, ["compress-ont", "-q", "4-fix", "-T", "5", "12", "20"]),
    "q_hifi_org": (dict(n_reads=200, genome_len=100000, mean_len=6000, seed=13, profile="hifi", n_frac=0.0), ["compress-pbhifi", "-q", "org"]),
    "q_clr_org": (dict(n_reads=300, genome_len=30000, mean_len=3000, seed=12, profile="clr", n_frac=0.03), ["compress-pbraw", "-q", "org"]),
    "multi_ont": (dict(n_reads=3000, genome_len=400000, mean_len=4000, seed=21, profile="ont", n_frac=0.02), ["compress-ont"]),
    "multi_ont_bal": (dict(n_reads=3000, genome_len=400000, mean_len=4000, seed=21, profile="ont", n_frac=0.02), ["compress-ont", "-p", "balanced"]),
}


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "colord")
    path = os.path.join(ROOT, "tests", "golden", "streams.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    only = sys.argv[1:]
    for name, (gen, cli) in CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            s = synth.generate(**gen)
            fq = os.path.join(tmp, "in.fastq")
            s.write_fastq(fq)
            subprocess.run([exe, *cli, "-t", "4", fq, os.path.join(tmp, "x.colord")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
            parts = colord_archive.read_parts(os.path.join(tmp, "x.colord"))
        out[name] = dict(generator=gen, cli=cli, bases_sha1=hashlib.sha1(s.bases.tobytes()).hexdigest(),
                         streams={k: [[md, len(b), hashlib.sha1(b).hexdigest()] for md, b in parts[k]] for k in ("dna", "qual", "header") if k in parts})
        print(name, {k: (len(v), sum(x[1] for x in v)) for k, v in out[name]["streams"].items()})
    with open(path, "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
