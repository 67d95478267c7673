"""Header fixtures for the header-stream tests (run in the build container, where /root/reference exists).

headers.json.gz holds, per case, the header lines of the reference's own test inputs (test/*.fastq: ONT uuid headers, SRA
headers, PacBio CLR headers) and the size of the header stream the unmodified reference (oracle/_ref/colord, default header
mode) wrote for that file; plus one synthetic case (20 000 ONT-style headers, recipe below) with the reference's size for it.
Usage: python tests/golden/make_hdr_golden.py
"""
import gzip
import json
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "..", "..", "oracle", "_ref", "colord")
CASES = {"ont": ("M.bovis.fastq", "compress-ont"), "hifi": ("D.melanogaster.fastq", "compress-pbhifi"), "clr": ("A.thaliana.fastq", "compress-pbraw")}


def synth_headers(n=20000, seed=1):
    rng = np.random.default_rng(seed)
    return [b"@read_%d ch=%d start_time=2020-01-01T00:%02d:%02dZ" % (i, int(rng.integers(1, 513)), (i // 60) % 60, i % 60) for i in range(n)]


def ref_header_size(fastq, mode):
    with tempfile.TemporaryDirectory() as d:
        out = subprocess.run([REF_BIN, mode, "-t", "4", fastq, os.path.join(d, "a.colord")], capture_output=True, text=True, cwd=d)
        return int(re.search(r"Header size\s*:\s*(\d+)", out.stdout + out.stderr).group(1))


def main():
    res = {}
    for name, (f, mode) in CASES.items():
        path = os.path.join("/root/reference/test", f)
        lines = open(path, "rb").read().split(b"\n")
        hs = [h.decode("latin-1") for h in lines[0::4] if h]
        res[name] = {"source": "test/" + f, "mode": mode, "headers": hs, "ref_header_stream_bytes": ref_header_size(path, mode)}
    hs = synth_headers()
    with tempfile.TemporaryDirectory() as d:
        rng = np.random.default_rng(7)
        p = os.path.join(d, "s.fastq")
        with open(p, "wb") as f:
            for h in hs:
                s = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 300))
                f.write(h + b"\n" + s + b"\n+\n" + b"5" * 300 + b"\n")
        res["synthetic_ont_20000"] = {"source": "make_hdr_golden.synth_headers(20000, 1)", "mode": "compress-ont", "headers": None,
                                      "ref_header_stream_bytes": ref_header_size(p, "compress-ont")}
    with gzip.open(os.path.join(HERE, "headers.json.gz"), "wt") as f:
        json.dump(res, f)
    print({k: v["ref_header_stream_bytes"] for k, v in res.items()})


if __name__ == "__main__":
    main()
