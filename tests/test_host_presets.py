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


CLI = os.path.join(ROOT, "colord_b200", "colord-b200")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "colord")


def _cli_params_block(args, cwd):
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "colord_b200", "csrc")], check=True, capture_output=True)
    r = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True)
    return r.stderr.split("\n")


@pytest.mark.parametrize("args", [["compress-ont"], ["compress-pbraw", "-p", "ratio"], ["compress-pbhifi", "-p", "balanced"], ["compress-ont", "-q", "none", "-p", "balanced"],
                                  ["compress-pbhifi", "-q", "org", "-k", "22", "-a", "19", "-c", "3"]])
def test_verbose_parameter_block(tmp_path, args):
    """`colord-b200 compress-* -v` prints the reference's parameter block (PrintParams, compression.cpp:165-207) line for line.
    The block is printed before the device is touched, so this runs without a GPU (the command then stops at clb_create)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("on a GPU box the command goes on to compress; the block is the same code")
    open(str(tmp_path / "s.fastq"), "w").write("@r1\n" + "ACGT" * 30 + "\n+\n" + "I" * 120 + "\n")
    ours = _cli_params_block(args + ["-v", "-t", "3", "s.fastq", "o.colord"], str(tmp_path))[:29]
    assert ours[0] == "input is not gzipped" and ours[1].startswith(" * * *") and "\tnumber of threads: 3" in ours
    key = " ".join(args[:1] + (["-p", args[args.index("-p") + 1]] if "-p" in args else ["-p", "memory"]))
    gold = GOLD["presets"].get(key)
    if gold and "-q" not in args and "-k" not in args:
        assert f"\tfilter modulo: {gold['filterHashModulo']}" in ours and f"\tmax k-mer count: {gold['maxKmerCount']}" in ours
        assert f"\tquality thresholds: {(' ' + gold['qualityFwdThresholds']) if gold['qualityFwdThresholds'] else ''}" in ours
    if os.path.exists(REF_BIN):
        ref = subprocess.run([REF_BIN] + args + ["-v", "-t", "3", "s.fastq", "o.colord"], cwd=str(tmp_path), capture_output=True, text=True)
        assert ours == ref.stderr.split("\n")[:29]
