#!/usr/bin/env python
"""Regenerate tests/golden/* from the UNMODIFIED reference (oracle/_ref/ref_stage_dump).

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle ref
drivers`).  For every case: generate the deterministic synthetic FASTQ (colord_b200.synth), run the
tapped reference on it with the case's CLI flags, gzip the per-stage dumps into tests/golden/<case>/ and
record the generator arguments + a SHA-1 of the bases so tests can regenerate the identical input on
the GPU box (where neither /root/reference nor the FASTQ exist).
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from colord_b200 import synth  # noqa: E402

CASES = {
    # name: (generator kwargs, reference CLI)
    "ont_mem": (dict(n_reads=400, genome_len=30000, mean_len=3000, seed=11, profile="ont", n_frac=0.05),
                ["compress-ont"]),
    "ont_bal": (dict(n_reads=400, genome_len=30000, mean_len=3000, seed=11, profile="ont", n_frac=0.05),
                ["compress-ont", "-p", "balanced"]),
    "clr_ratio": (dict(n_reads=300, genome_len=30000, mean_len=3000, seed=12, profile="clr", n_frac=0.03),
                  ["compress-pbraw", "-p", "ratio"]),
    "hifi": (dict(n_reads=200, genome_len=100000, mean_len=6000, seed=13, profile="hifi", n_frac=0.0),
             ["compress-pbhifi"]),
}


def main():
    dumper = os.path.join(ROOT, "oracle", "_ref", "ref_stage_dump")
    if not os.path.exists(dumper):
        sys.exit("build the oracle first: make -C oracle ref drivers")
    only = sys.argv[1:]
    for name, (gen, cli) in CASES.items():
        if only and name not in only:
            continue
        out = os.path.join(ROOT, "tests", "golden", name)
        os.makedirs(out, exist_ok=True)
        with tempfile.TemporaryDirectory() as tmp:
            s = synth.generate(**gen)
            fq = os.path.join(tmp, "in.fastq")
            s.write_fastq(fq)
            env = dict(os.environ, COLORD_DUMP_DIR=os.path.join(tmp, "dump"))
            subprocess.run([dumper, *cli, "-t", "4", fq, os.path.join(tmp, "x.out")], check=True, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
            for f in ("kmers.bin", "reads.bin", "es.bin"):
                with open(os.path.join(tmp, "dump", f), "rb") as src, \
                        gzip.GzipFile(os.path.join(out, f + ".gz"), "wb", compresslevel=9, mtime=0) as dst:
                    shutil.copyfileobj(src, dst)
            shutil.copy(os.path.join(tmp, "dump", "params.txt"), os.path.join(out, "params.txt"))
            meta = dict(generator=gen, cli=cli, n_reads=s.n_reads, n_bases=s.n_bases,
                        bases_sha1=hashlib.sha1(s.bases.tobytes()).hexdigest(),
                        offsets_sha1=hashlib.sha1(s.offsets.tobytes()).hexdigest())
            with open(os.path.join(out, "input.json"), "w") as f:
                json.dump(meta, f, indent=1)
        print(name, {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})


if __name__ == "__main__":
    main()
