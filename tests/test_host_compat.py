"""`colord-b200 decompress` on archives written by the UNMODIFIED reference: the host decoders of the reference's own streams
(colord_b200/host/compat_decoder.h) must print what the reference's `decompress` prints.  CPU only (decompression is host code).
Fixtures: tests/golden/ref_archives (generator make_ref_archive_golden.py: one archive per stream flavour + the SHA-1 of the reference's
output); where the stock binary is present (oracle/_ref/colord, built by oracle/Makefile) also a live comparison on an input of
several packs — the coders restart per pack while their models live on."""
import hashlib
import json
import os
import subprocess

import pytest

from colord_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colord_b200", "colord-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "colord")
DIR = os.path.join(ROOT, "tests", "golden", "ref_archives")
with open(os.path.join(DIR, "expected.json")) as f:
    EXPECTED = json.load(f)

pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="colord-b200 is not built")


def _genome_args(name, tmp_path):
    """-G archives without -s need the genome again: rebuilt as the fixture generator wrote it"""
    e = EXPECTED[name]
    if "-G" not in e["cli"] or "-s" in e["cli"]:
        return []
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_ref_archive_golden.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    p = str(tmp_path / "genome.fa")
    mk.write_genome(p, e["generator"])
    return ["-G", p]


@pytest.mark.parametrize("name", list(EXPECTED))
def test_reference_archive_decodes_like_the_reference(name, tmp_path):
    out = str(tmp_path / "out")
    r = subprocess.run([CLI, "decompress", *_genome_args(name, tmp_path), os.path.join(DIR, name + ".colord"), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    data = open(out, "rb").read()
    e = EXPECTED[name]
    assert len(data) == e["output_bytes"]
    assert hashlib.sha1(data).hexdigest() == e["output_sha1"]
    if e["lossless"]:
        assert hashlib.sha1(data).hexdigest() == e["input_sha1"]


def test_genome_archives_refuse_a_missing_or_different_genome(tmp_path):
    arc = os.path.join(DIR, "ont_genome_checksum_bal.colord")
    r = subprocess.run([CLI, "decompress", arc, str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "reference genome is required" in r.stderr
    g = str(tmp_path / "other.fa")
    open(g, "w").write(">x\nACGTACGTTTGACCA\n")
    r = subprocess.run([CLI, "decompress", "-G", g, arc, str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "different reference genome" in r.stderr


def test_info_of_a_reference_archive():
    r = subprocess.run([CLI, "info", os.path.join(DIR, "ont_default.colord")], capture_output=True, text=True)
    assert r.returncode == 0 and "version major: 1" in r.stderr and "total reads: 120" in r.stderr


@pytest.mark.parametrize("damage", ["flip", "cut", "zero"])
def test_damaged_reference_archives_end_in_an_error(damage, tmp_path):
    """every part of a stream damaged in turn: exit code 0 (still decodable) or 1 (refused), never a signal"""
    src = open(os.path.join(DIR, "ont_balanced_5avg.colord"), "rb").read()
    for k in range(12):
        b = bytearray(src)
        at = 8 + (len(b) - 400) * k // 12
        if damage == "flip":
            b[at] ^= 0x5a
        elif damage == "zero":
            b[at:at + 64] = bytes(64)
        else:
            b = b[:at] + b[-300:]
        p = str(tmp_path / f"d{k}.colord")
        open(p, "wb").write(bytes(b))
        r = subprocess.run([CLI, "decompress", p, str(tmp_path / "o")], capture_output=True)
        assert r.returncode in (0, 1), (damage, k, r.returncode, r.stderr[-200:])


@pytest.mark.skipif(not os.path.exists(REF), reason="the stock reference binary is not built here")
@pytest.mark.parametrize("cli", [["compress-ont"], ["compress-ont", "-p", "balanced", "-q", "org"]])
def test_live_against_the_stock_binary_three_packs(cli, tmp_path):
    fq = str(tmp_path / "in.fastq")
    synth.generate_file(fq, "ont", 1300, 500_000, 8000, seed=5, workers=4)
    arc, a, b = str(tmp_path / "x.colord"), str(tmp_path / "ref.out"), str(tmp_path / "b200.out")
    subprocess.run([REF, *cli, "-t", "4", fq, arc], check=True, capture_output=True, cwd=str(tmp_path))
    subprocess.run([REF, "decompress", arc, a], check=True, capture_output=True, cwd=str(tmp_path))
    r = subprocess.run([CLI, "decompress", arc, b], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run(["cmp", "-s", a, b]).returncode == 0
