"""Parameter fixtures for tests/test_host_presets.py (run in the build container, where /root/reference exists).

presets.json holds, for each compression mode x priority (and a few -q modes), the parameter block the unmodified reference
prints under -v (PrintParams, src/colord/compression.cpp:165-207) on a two-read input, plus the derived values it reports
(tot k-mers, n uniq counted, approx. avg. read len, sparse mode range in reads) for the reference's three test files.
Usage: python tests/golden/make_presets_golden.py
"""
import json
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.abspath(os.path.join(HERE, "..", "..", "oracle", "_ref", "colord"))
T = "/root/reference/test"
KEYS = {"k-mer length": "kmerLen", "anchor length": "anchorLen", "data source type": "dataSource", "filter modulo": "filterHashModulo", "max candidates": "maxCandidates",
        "min k-mer count": "minKmerCount", "max k-mer count": "maxKmerCount", "max matches multiplier": "maxMatchesMultiplier", "max recurence": "maxRecurence",
        "min anchors": "minAnchors", "min fraction of m-mers in encode": "minFractionOfMmersInEncode",
        "min fraction of m-mers in encode to always encode": "minFractionOfMmersInEncodeToAlwaysEncode",
        "min part length to consider alternative reference read": "minPartLenToConsiderAltRead", "compression priority": "priority",
        "quality compression mode": "qualityComprMode", "quality thresholds": "qualityFwdThresholds", "quality values": "qualityRevThresholds",
        "reference reads mode": "referenceReadsMode", "sparse mode exponent": "sparseMode_exponent", "sparse mode range": "sparseMode_range_symbols",
        "multipier for predicted cost of storing read part as edit script": "editScriptCostMultiplier", "header compression mode": "headerComprMode"}


def run(args, cwd):
    out = subprocess.run([REF_BIN] + args, cwd=cwd, capture_output=True, text=True)
    return out.stdout + out.stderr


def params_block(text):
    res = {}
    for line in text.split("\n"):
        m = re.match(r"\t([^:]+): ?(.*)$", line)
        if m and m.group(1) in KEYS:
            res[KEYS[m.group(1)]] = m.group(2).strip()
    return res


def main():
    res = {"presets": {}, "derived": {}}
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "s.fastq"), "w") as f:
            f.write("\n".join(open(os.path.join(T, "M.bovis.fastq")).read().split("\n")[:8]) + "\n")
        for mode in ("compress-ont", "compress-pbraw", "compress-pbhifi"):
            for pri in ("ratio", "balanced", "memory"):
                res["presets"][f"{mode} -p {pri}"] = params_block(run([mode, "-p", pri, "-v", "-t", "2", "s.fastq", "o.colord"], d))
        for q in ("org", "none", "2-avg", "4-avg", "5-avg", "2-fix", "4-fix", "5-fix", "avg"):
            res["presets"][f"compress-ont -q {q}"] = params_block(run(["compress-ont", "-q", q, "-v", "-t", "2", "s.fastq", "o.colord"], d))
        for name, mode, extra in (("M.bovis.fastq", "compress-ont", []), ("M.bovis.fastq", "compress-ont", ["-p", "balanced"]), ("D.melanogaster.fastq", "compress-pbhifi", []),
                                  ("A.thaliana.fastq", "compress-pbraw", ["-p", "balanced"])):
            text = run([mode] + extra + ["-v", "-t", "4", os.path.join(T, name), "o.colord"], d)
            g = lambda pat: int(re.search(pat, text).group(1))
            p = params_block(text)
            res["derived"][" ".join([mode] + extra + [name])] = {
                "file_bytes": os.path.getsize(os.path.join(T, name)), "kmerLen": int(p["kmerLen"]), "anchorLen": int(p["anchorLen"]),
                "filterHashModulo": int(p["filterHashModulo"]), "sparseMode_range_symbols": float(p["sparseMode_range_symbols"]),
                "tot_kmers": g(r"tot k-mers: (\d+)"), "n_uniq_counted": g(r"n uniq counted: (\d+)"), "mean_read_len": g(r"approx\. avg\. read len: (\d+)"),
                "sparse_range_reads": g(r"sparse mode range in reads: (\d+)"), "n_reads": len(open(os.path.join(T, name)).read().split("\n")) // 4}
    json.dump(res, open(os.path.join(HERE, "presets.json"), "w"), indent=1, sort_keys=True)
    print(len(res["presets"]), "presets,", len(res["derived"]), "derived")


if __name__ == "__main__":
    main()
