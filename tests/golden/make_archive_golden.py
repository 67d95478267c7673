"""Archive fixtures for tests/test_host_archive.py (run in the build container, where /root/reference exists).

tests/golden/archives/<case>.colord are archives written by the unmodified reference (oracle/_ref/colord) from the first reads
of its own test files, one per shape of the `meta` record (quality mode with 0 / 1 / 4 thresholds, sparse / all reference
reads, FASTA input without a quality stream, reference genome stored / given by checksum).  expected.json holds, per case,
what `colord info` of the reference prints for the archive and the command that made it.
Usage: python tests/golden/make_archive_golden.py
"""
import json
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.abspath(os.path.join(HERE, "..", "..", "oracle", "_ref", "colord"))
OUT = os.path.join(HERE, "archives")
T = "/root/reference/test"

# case -> (input file, reads kept, mode, extra arguments)
CASES = {
    "ont_default": ("M.bovis.fastq", 6, "compress-ont", []),
    "hifi_org": ("D.melanogaster.fastq", 3, "compress-pbhifi", ["-q", "org"]),
    "clr_none_all": ("A.thaliana.fastq", 4, "compress-pbraw", ["-q", "none", "-p", "ratio"]),
    "ont_4fix": ("M.bovis.fastq", 6, "compress-ont", ["-q", "4-fix", "-p", "balanced"]),
    "ont_fasta": ("M.bovis.fastq", 6, "compress-ont", []),            # converted to FASTA: no qual stream
    "ont_genome_checksum": ("M.bovis.fastq", 6, "compress-ont", ["-G", "genome.fa"]),
    "ont_genome_stored": ("M.bovis.fastq", 6, "compress-ont", ["-G", "genome.fa", "-s"]),
}


def parse_info(text):
    g = lambda k: re.search(rf"{k}: (.*)", text).group(1).strip()
    return {"version": [int(g("version major")), int(g("version minor")), int(g("version patch"))], "total_bytes": int(g("total bytes")),
            "total_bases": int(g("total bases")), "total_reads": int(g("total reads")), "command": g("command")}


def main():
    os.makedirs(OUT, exist_ok=True)
    exp = {}
    with tempfile.TemporaryDirectory() as d:
        # a 60 kb piece of the reference's test genome
        with open(os.path.join(T, "M.bovis-reference.fna")) as f, open(os.path.join(d, "genome.fa"), "w") as g:
            for i, line in enumerate(f):
                if i > 750:
                    break
                g.write(line)
        for name, (src, n_reads, mode, extra) in CASES.items():
            lines = open(os.path.join(T, src)).read().split("\n")[:4 * n_reads]
            inp = "in.fastq"
            if name == "ont_fasta":
                inp = "in.fasta"
                lines = [x for r in range(n_reads) for x in (">" + lines[4 * r][1:], lines[4 * r + 1])]
            with open(os.path.join(d, inp), "w") as f:
                f.write("\n".join(lines) + "\n")
            cmd = ["colord", mode, "-t", "2"] + extra + [inp, name + ".colord"]
            subprocess.run([REF_BIN] + cmd[1:], cwd=d, check=True, capture_output=True)
            info = subprocess.run([REF_BIN, "info", name + ".colord"], cwd=d, check=True, capture_output=True, text=True)
            data = open(os.path.join(d, name + ".colord"), "rb").read()
            open(os.path.join(OUT, name + ".colord"), "wb").write(data)
            exp[name] = {"made_by": " ".join(cmd), "bytes": len(data), "info": parse_info(info.stdout + info.stderr)}
            print(name, len(data), "bytes")
    json.dump(exp, open(os.path.join(OUT, "expected.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
