"""Host-side archive container + `meta` / `info` records against archives written by the unmodified
reference (tests/golden/archives, made by make_archive_golden.py).  Host code only: no device, no oracle."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCH = os.path.join(ROOT, "tests", "golden", "archives")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "colord")
EXPECTED = json.load(open(os.path.join(ARCH, "expected.json")))
CASES = sorted(EXPECTED)


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_archive") / "tool")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", out, os.path.join(ROOT, "tests", "host_archive_tool.cpp")], check=True)
    return out


@pytest.mark.parametrize("case", CASES)
def test_rewrite_is_byte_identical(tool, tmp_path, case):
    """Reading every part of a reference archive and writing it back through the writer reproduces the file."""
    src = os.path.join(ARCH, case + ".colord")
    dst = str(tmp_path / "copy.colord")
    subprocess.run([tool, "rewrite", src, dst], check=True)
    assert open(dst, "rb").read() == open(src, "rb").read()
    assert os.path.getsize(src) == EXPECTED[case]["bytes"]


@pytest.mark.parametrize("case", CASES)
def test_meta_and_info_records(tool, case):
    """`meta` and `info` parse to the values the reference prints / was run with, and serialise back to the same bytes."""
    src = os.path.join(ARCH, case + ".colord")
    assert subprocess.run([tool, "meta-roundtrip", src]).returncode == 0
    d = json.loads(subprocess.run([tool, "dump", src], check=True, capture_output=True, text=True).stdout)
    want = EXPECTED[case]["info"]
    got = d["info"]
    assert got["version"] == want["version"] and got["total_bytes"] == want["total_bytes"]
    assert got["total_bases"] == want["total_bases"] and got["total_reads"] == want["total_reads"]
    # the stored command line carries the path of the binary; `colord info` prints it as stored
    assert got["command"] == want["command"]
    m = d["meta"]
    names = [s["name"] for s in d["streams"]]
    cmd = EXPECTED[case]["made_by"].split()
    assert names[0] == "meta" and names[-1] == "info" and "dna" in names and "header" in names
    assert m["is_fastq"] == (1 if "qual" in names else 0) == (0 if case == "ont_fasta" else 1)
    assert m["dataSource"] == {"compress-ont": 0, "compress-pbraw": 1, "compress-pbhifi": 2}[cmd[1]]
    if m["is_fastq"]:
        q = cmd[cmd.index("-q") + 1] if "-q" in cmd else {"compress-ont": "4-avg", "compress-pbhifi": "5-avg", "compress-pbraw": "none"}[cmd[1]]
        assert m["qualityComprMode"] == {"org": 0, "5-avg": 1, "4-avg": 2, "2-avg": 3, "5-fix": 4, "4-fix": 5, "2-fix": 6, "avg": 7, "none": 8}[q]
        assert len(m["qualityRevThresholds"]) == {"none": 1, "2-fix": 2, "4-fix": 4, "5-fix": 5}.get(q, 0)
    pri = cmd[cmd.index("-p") + 1] if "-p" in cmd else "memory"
    # arg_parse.cpp:95-385: level 3 / 2 / 1 for ratio / balanced / memory, except pbhifi memory = 2
    assert m["compressionLevel"] == {"memory": 2 if cmd[1] == "compress-pbhifi" else 1, "balanced": 2, "ratio": 3}[pri]
    assert m["referenceReadsMode"] == (0 if pri == "ratio" else 1)            # ratio keeps all reads as references
    assert m["ref_genome_available"] == (1 if "-G" in cmd else 0)
    assert m["storeRefGenome"] == (1 if "-s" in cmd else 0)
    assert len(m["ref_genome_checksum"]) == (32 if "-G" in cmd and "-s" not in cmd else 0)
    assert ("ref-genome" in names) == ("-s" in cmd)
    # part tables: every part lies inside the file, parts do not overlap
    spans = sorted((p[0], p[1]) for s in d["streams"] for p in s["parts"])
    for (o0, n0), (o1, _) in zip(spans, spans[1:]):
        assert o0 + n0 < o1 + 1
    assert spans[-1][0] + spans[-1][1] < EXPECTED[case]["bytes"]


def test_archive_from_scratch(tool, tmp_path):
    """An archive written from scratch (slots filled out of order, an empty part, every optional meta field) reads back."""
    p = str(tmp_path / "made.colord")
    subprocess.run([tool, "make", p], check=True)
    d = json.loads(subprocess.run([tool, "dump", p], check=True, capture_output=True, text=True).stdout)
    dna = next(s for s in d["streams"] if s["name"] == "dna")
    assert dna["raw_size"] == 99 and [x[1:] for x in dna["parts"]] == [[300, 0], [5, 1234567]]
    assert dna["parts"][1][0] == 0 and dna["parts"][0][0] > dna["parts"][1][0]          # the second slot was written first
    assert next(s for s in d["streams"] if s["name"] == "qual")["parts"][0][1:] == [0, 0]
    m = d["meta"]
    assert m["qualityRevThresholds"] == [3, 10, 18, 35] and m["sparseMode_exponent"] == 1.25 and m["ref_genome_checksum"] == "ab" * 16
    assert subprocess.run([tool, "meta-roundtrip", p]).returncode == 0
    # rewriting our own archive is stable too
    q = str(tmp_path / "made2.colord")
    subprocess.run([tool, "rewrite", p, q], check=True)
    assert open(p, "rb").read() == open(q, "rb").read()


def test_damaged_archive_is_refused(tool, tmp_path):
    src = os.path.join(ARCH, "ont_default.colord")
    data = open(src, "rb").read()
    for name, blob in {"truncated": data[:-9], "empty": b"", "footer_too_long": data[:-8] + (1 << 40).to_bytes(8, "little")}.items():
        p = str(tmp_path / (name + ".colord"))
        open(p, "wb").write(blob)
        assert subprocess.run([tool, "dump", p], capture_output=True).returncode == 2, name


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs the reference binary (oracle/_ref, build container only)")
def test_reference_reads_our_archive(tool, tmp_path):
    """The unmodified reference opens an archive written by the host-side writer and prints its info record."""
    p = str(tmp_path / "made.colord")
    subprocess.run([tool, "make", p], check=True)
    out = subprocess.run([REF_BIN, "info", p], capture_output=True, text=True)
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "total bases: 4" in text and "command: made by host_archive_tool" in text and "version minor: 2" in text
    # and a rewritten reference archive still decompresses with the reference
    src = os.path.join(ARCH, "ont_default.colord")
    dst = str(tmp_path / "copy.colord")
    subprocess.run([tool, "rewrite", src, dst], check=True)
    a, b = str(tmp_path / "a.fastq"), str(tmp_path / "b.fastq")
    subprocess.run([REF_BIN, "decompress", src, a], check=True, capture_output=True)
    subprocess.run([REF_BIN, "decompress", dst, b], check=True, capture_output=True)
    assert open(a, "rb").read() == open(b, "rb").read() and os.path.getsize(a) > 0
