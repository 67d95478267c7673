"""Quality-stream goldens from the reference's own test fixtures (run in the build container; /root/reference is read-only).

test/M.bovis.fastq(.quan)        ONT default  (compress-ont,    4-avg, thresholds 7 14 26)    -- .github/workflows/main.yml:34-38
test/D.melanogaster.fastq(.quan) HiFi default (compress-pbhifi, 5-avg, thresholds 7 14 26 93) -- main.yml:40-44
The .quan files are what `colord decompress` prints for the lossy modes; they pin the reference's quantisation, per-bin means
and error-diffusion reconstruction.  Stored: bases, original qualities, expected qualities of the first N reads.
"""
import gzip
import struct
import sys

import numpy as np

REF = "/root/reference/test/"


def read_fastq(p, n_max):
    lines = open(p, "rb").read().split(b"\n")
    seqs = [s for s in lines[1::4] if s][:n_max]
    quals = lines[3::4][:len(seqs)]
    return seqs, quals


def make(name, fastq, n_max, out):
    seqs, quals = read_fastq(REF + fastq, n_max)
    seqs2, quan = read_fastq(REF + fastq + ".quan", n_max)
    assert seqs == seqs2
    with gzip.open(out, "wb", 9) as f:
        f.write(struct.pack("<I", len(seqs)))
        f.write(np.array([len(s) for s in seqs], np.uint32).tobytes())
        f.write(b"".join(seqs)); f.write(b"".join(quals)); f.write(b"".join(quan))
    print(name, len(seqs), "reads", sum(map(len, seqs)), "bases ->", out)


if __name__ == "__main__":
    d = sys.argv[1] if len(sys.argv) > 1 else "."
    make("ont", "M.bovis.fastq", 100, d + "/qual_ont.bin.gz")
    make("hifi", "D.melanogaster.fastq", 40, d + "/qual_hifi.bin.gz")
