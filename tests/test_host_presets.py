"""Parameter defaults and derived values (colord_b200/host/presets.h) against what the unmodified reference prints under -v
(tests/golden/presets.json, written by make_presets_golden.py from oracle/_ref/colord).  Host code only."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "presets.json")))


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_presets") / "tool")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", out, os.path.join(ROOT, "tests", "host_presets_tool.cpp")], check=True)
    return out


@pytest.mark.parametrize("case", sorted(GOLD["presets"]))
def test_default_sets(tool, case):
    """Every number of the nine default sets and of the -q modes' thresholds equals the reference's (arg_parse.cpp:28-376)."""
    want = GOLD["presets"][case]
    words = case.split()
    mode, pri, q = words[0], "memory", None
    if "-p" in words:
        pri = words[words.index("-p") + 1]
    if "-q" in words:
        q = words[words.index("-q") + 1]
    got = json.loads(subprocess.run([tool, "preset", mode, pri] + ([q] if q else []), check=True, capture_output=True, text=True).stdout)
    for key, val in want.items():
        if key in ("kmerLen", "anchorLen"):
            continue                      # derived from the input size: test_derived_values
        assert str(got[key]) == val, (case, key, got[key], val)
    # the archive's compression level follows the priority (arg_parse.cpp:95-385)
    assert got["compressionLevel"] == {"ratio": 3, "balanced": 2, "memory": 2 if mode == "compress-pbhifi" else 1}[pri]


@pytest.mark.parametrize("case", sorted(GOLD["derived"]))
def test_derived_values(tool, case):
    """k / anchor length from the file size, mean read length and sparse range from the k-mer statistics (compression.cpp:41-93, :443, :501-504)."""
    w = GOLD["derived"][case]
    got = json.loads(subprocess.run([tool, "derive", str(w["file_bytes"]), "0", "1", str(w["tot_kmers"]), str(w["filterHashModulo"]), str(w["n_reads"]),
                                     str(w["n_uniq_counted"]), str(w["sparseMode_range_symbols"])], check=True, capture_output=True, text=True).stdout)
    for key in ("kmerLen", "anchorLen", "mean_read_len", "sparse_range_reads"):
        assert got[key] == w[key], (case, key, got[key], w[key])


@pytest.mark.parametrize("file_bytes,gz,fastq,k,a", [(2_000_000_000, 0, 1, 20, 16), (2_100_000_000, 0, 1, 21, 18), (9_000_000_000, 0, 1, 23, 21), (50_000_000_000, 0, 1, 24, 22),
                                                     (99_000_000_000, 0, 1, 25, 22), (300_000_000_000, 0, 1, 26, 23), (500_000_000, 1, 1, 21, 18), (5_000_000_000, 0, 0, 23, 21),
                                                     (300_000_000, 1, 0, 21, 18)])
def test_kmer_length_table(tool, file_bytes, gz, fastq, k, a):
    """compression.cpp:52-92: bases ~ 0.49 x bytes (FASTQ), 0.98 (FASTA), 2.08 / 3.98 when gzipped; 50 GB FASTQ -> k 24, anchors 22."""
    got = json.loads(subprocess.run([tool, "derive", str(file_bytes), str(gz), str(fastq), "1000", "12", "10", "10", "1"], check=True, capture_output=True, text=True).stdout)
    assert (got["kmerLen"], got["anchorLen"]) == (k, a)
