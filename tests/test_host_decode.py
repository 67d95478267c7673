"""Host-side decoders of the native containers (colord_b200/host/decompressor.h) and the input reader (fastq_reader.h) on CPU.

The containers decoded here are written by the oracle's CPU twins of the device encoders (byte-identical to the device's output,
tests/test_gpu_stage3.py) or are device-made fixtures (tests/golden/b200_archives, written on a B200 by make_b200_archives.py).
The product decoders share no code with the oracle: equal results check both.  The oracle is used as the checker only."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import golden_io
import oracle_lib
from colord_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "colord")      # the unmodified reference (oracle/Makefile), built where /root/reference exists


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_decode") / "tool")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-o", out, os.path.join(ROOT, "tests", "host_decode_tool.cpp"), "-lz", "-pthread"], check=True)
    return out


def _write(path, arr):
    np.ascontiguousarray(arr).tofile(path)
    return path


def _packs(offsets):
    """the reference's read-pack rule (in_reads.cpp:62-76): a pack closes at >= 4 MiB counting a guard byte per read"""
    packs, cur, n = [], 0, 0
    for i in range(len(offsets) - 1):
        cur += int(offsets[i + 1] - offsets[i]) + 1
        n += 1
        if cur >= (2 << 21):
            packs.append(n)
            cur, n = 0, 0
    if n:
        packs.append(n)
    return packs


# ---------------------------------------------------------------------------------------------------------------- headers
@pytest.mark.parametrize("case", ["ont", "hifi", "clr", "synthetic_ont_20000"])
def test_header_decoder(tool, tmp_path, case):
    headers, _ = golden_io.load_hdr_golden()[case]
    headers = headers[:6000]
    plus = (np.arange(len(headers)) % 3 == 0).astype(np.uint8)
    stream = oracle_lib.hdr_encode(headers, plus=plus, packs=[min(1500, len(headers) - i) for i in range(0, len(headers), 1500)])
    s, o = _write(str(tmp_path / "s"), stream), str(tmp_path / "o")
    subprocess.run([tool, "hdr", s, str(len(headers)), o], check=True)
    got = open(o, "rb").read().split(b"\n")[:-1]
    assert len(got) == len(headers)
    for g, h, p in zip(got, headers, plus):
        assert g == b"%d\t" % p + h


# -------------------------------------------------------------------------------------------------------------- qualities
@pytest.mark.parametrize("name,n_bins,thr", [("ont", 4, [7, 14, 26]), ("hifi", 5, [7, 14, 26, 93]), ("ont", 2, [7])])
def test_quality_avg_decoder_reproduces_the_reference_quan(tool, tmp_path, name, n_bins, thr):
    """QB01 (level 1) -> the qualities `colord decompress` of the reference prints (its own .quan fixtures for 4-avg / 5-avg)."""
    bases, quals, quan, off = golden_io.load_qual_golden(name)
    P = oracle_lib.qual_params(n_bins, thr, 1)
    stream = oracle_lib.qual_encode(P, bases, quals, off, _packs(off))
    s, b, f, o = _write(str(tmp_path / "s"), stream), _write(str(tmp_path / "b"), bases), _write(str(tmp_path / "f"), off), str(tmp_path / "o")
    subprocess.run([tool, "qavg", s, b, f, o], check=True)
    got = np.fromfile(o, np.uint8)
    if n_bins != 2:
        assert np.array_equal(got, quan)            # the reference's own fixture
    assert np.array_equal(got, oracle_lib.qual_lossy(P, bases, quals, off))


def _fake_tuples(rng, lens):
    """CompactES-shaped tuples whose match / anchor / other layout is random (flags only depend on the tuple types)"""
    es, es_off, flags = [], [0], []
    for n in lens:
        out = bytearray([10 << 4, 0, 0, 0, 0])
        fl, at = [], 0
        while at < n:
            t = int(rng.integers(0, 5))
            if t == 4 and n - at >= 15:
                ln = int(rng.integers(15, min(60, n - at) + 1))
                out += bytes([(4 << 4) | (ln >> 24), (ln >> 16) & 255, (ln >> 8) & 255, ln & 255])
                fl += [2] * ln
                at += ln
            elif t == 1:
                out.append(1 << 4)              # deletion: no base
            elif t == 2:
                out.append(2 << 4); fl.append(1); at += 1
            else:
                out.append((0 if t == 0 else 3) << 4); fl.append(0); at += 1
        es.append(bytes(out)); es_off.append(es_off[-1] + len(out)); flags += fl
    return np.frombuffer(b"".join(es), np.uint8).copy(), np.array(es_off, np.uint64), np.array(flags, np.uint8)


@pytest.mark.parametrize("level", [2, 3])
def test_quality_avg_decoder_with_tuple_flags(tool, tmp_path, level):
    s_ = synth.generate(60, 20000, 1500, seed=5, profile="ont")
    rng = np.random.default_rng(7)
    lens = np.diff(s_.offsets).astype(np.int64)
    es, es_off, flags = _fake_tuples(rng, lens)
    P = oracle_lib.qual_params(4, [7, 14, 26], level)
    stream = oracle_lib.qual_encode(P, s_.bases, s_.quals, s_.offsets, [40, 20], es=es, es_off=es_off)
    paths = [_write(str(tmp_path / n), a) for n, a in (("s", stream), ("b", s_.bases), ("f", s_.offsets), ("fl", flags))]
    o = str(tmp_path / "o")
    subprocess.run([tool, "qavg", *paths, o], check=True)
    assert np.array_equal(np.fromfile(o, np.uint8), oracle_lib.qual_lossy(P, s_.bases, s_.quals, s_.offsets))


@pytest.mark.parametrize("source,level", [(0, 1), (2, 2), (1, 3), (2, 3)])
def test_quality_org_decoder_is_lossless(tool, tmp_path, source, level):
    s_ = synth.generate(50, 20000, 1500, seed=11 + source, profile="hifi" if source == 2 else "ont")
    rng = np.random.default_rng(3)
    es, es_off, flags = _fake_tuples(rng, np.diff(s_.offsets).astype(np.int64))
    quals = s_.quals.copy()
    quals[::97] = 33 + 93                                                    # the top value has its own quantiser cell
    stream = oracle_lib.qorg_encode(source, level, s_.bases, quals, s_.offsets, [30, 20], es=es if level > 1 else None, es_off=es_off if level > 1 else None)
    paths = [_write(str(tmp_path / n), a) for n, a in (("s", stream), ("b", s_.bases), ("f", s_.offsets), ("fl", flags))]
    o = str(tmp_path / "o")
    subprocess.run([tool, "qorg", *paths, o], check=True)
    assert np.array_equal(np.fromfile(o, np.uint8), quals)


def test_damaged_streams_are_refused(tool, tmp_path):
    headers, _ = golden_io.load_hdr_golden()["ont"]
    stream = oracle_lib.hdr_encode(headers[:50])
    for name, blob in {"cut": stream[:len(stream) // 2], "magic": np.concatenate([np.frombuffer(b"XX01", np.uint8), stream[4:]]), "empty": stream[:0]}.items():
        s = _write(str(tmp_path / name), blob)
        r = subprocess.run([tool, "hdr", s, "50", str(tmp_path / "o")], capture_output=True)
        assert r.returncode == 3, (name, r.returncode, r.stderr)
    s = _write(str(tmp_path / "count"), stream)
    assert subprocess.run([tool, "hdr", s, "51", str(tmp_path / "o")], capture_output=True).returncode == 3


def test_crafted_dna_stream_with_oversized_alphabet_is_refused(tool, tmp_path):
    """A DB01 header announcing max_cand = 1000 would make the short-id family's alphabet larger than the decoder's table scratch:
    refused at the header (the encoder's own limit is 32), and a script can never append past the archive's base count."""
    import struct
    blob = b"DB01" + struct.pack("<IIQII", 1, 1000, 1, 1, 0) + bytes(4096)
    s = _write(str(tmp_path / "crafted"), np.frombuffer(blob, np.uint8))
    d = _write(str(tmp_path / "dec"), np.ones(1, np.uint8))
    r = subprocess.run([tool, "dna", s, "1", d, str(tmp_path / "b"), str(tmp_path / "o"), str(tmp_path / "f")], capture_output=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    blob = b"DB01" + struct.pack("<IIQII", 1, 0, 1, 1, 0) + bytes(4096)
    s = _write(str(tmp_path / "crafted0"), np.frombuffer(blob, np.uint8))
    assert subprocess.run([tool, "dna", s, "1", d, str(tmp_path / "b"), str(tmp_path / "o"), str(tmp_path / "f")], capture_output=True).returncode == 3


# ------------------------------------------------------------------------------------------------------------------ reader
def _parse(tool, path, prefix):
    r = subprocess.run([tool, "parse", path, prefix], capture_output=True, text=True)
    return r, (json.loads(r.stdout) if r.returncode == 0 else None)


def _fastq(records, eol="\n", plus_header=()):
    out = []
    for i, (h, s, q) in enumerate(records):
        out += ["@" + h, s, "+" + (h if i in plus_header else ""), q]
    return (eol.join(out) + eol).encode()


RECORDS = [("r1 a=1", "ACGTNACGT", "IIIIIIIII"), ("r2", "TTTT", "!!!!"), ("r3/x", "GATTACA", "5555555")]


@pytest.mark.parametrize("threads,piece", [(1, 1 << 20), (3, 1 << 20), (8, 3 << 20)])
@pytest.mark.parametrize("eol,plus", [("\n", False), ("\r\n", True)])
def test_streaming_reader_equals_the_whole_file_reader(tool, tmp_path, threads, piece, eol, plus):
    """The streaming form (pieces cut at record starts, a ring of reused buffers, pieces handed over in file order) delivers exactly
    what the whole-file reader holds — bases, qualities, offsets, headers, '+' flags, statistics, pack sizes — also when quality
    lines begin with '@' or '+', reads are longer than a piece, and lines end in CRLF."""
    rng = np.random.default_rng(threads * 7 + piece)
    recs = []
    A, Q = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTN", np.uint8), np.frombuffer(b"@+!I5~", np.uint8)
    for i in range(600):
        n = int(rng.integers(1, 9000)) if i % 97 else int(2.5 * piece)
        recs.append((f"r{i} x={i * 7}", A[rng.integers(0, len(A), n)].tobytes().decode(), Q[rng.integers(0, len(Q), n)].tobytes().decode()))
    data = _fastq(recs, eol=eol, plus_header=range(0, 600, 3) if plus else ())
    p = str(tmp_path / "in.fastq")
    open(p, "wb").write(data)
    r1, st1 = _parse(tool, p, str(tmp_path / "a"))
    assert r1.returncode == 0, r1.stderr
    r2 = subprocess.run([tool, "stream", p, str(tmp_path / "b"), str(threads), str(piece)], capture_output=True, text=True)
    assert r2.returncode == 0, (r2.returncode, r2.stderr)
    st2 = json.loads(r2.stdout)
    for k in ("n_reads", "total_bytes", "total_bases", "total_symb_header", "read_packs", "header_packs"):
        assert st1[k] == st2[k], k
    for ext in ("bases", "quals", "offsets", "headers", "hoff", "plus"):
        assert open(str(tmp_path / f"a.{ext}"), "rb").read() == open(str(tmp_path / f"b.{ext}"), "rb").read(), ext


@pytest.mark.parametrize("what", ["bad_symbol", "no_final_eol", "quality_length", "plus_differs"])
def test_streaming_reader_hands_irregular_inputs_back(tool, tmp_path, what):
    """anything irregular ends the streaming form with a request for the whole-file reader (exit code 5 of the tool), which
    follows the reference line by line and words the refusal"""
    recs = [(f"r{i}", "ACGT" * 300, "IIII" * 300) for i in range(4000)]
    data = _fastq(recs)
    if what == "bad_symbol":
        data = data.replace(b"ACGTACGT", b"ACXTACGT", 1)
    elif what == "no_final_eol":
        data = data[:-1]
    elif what == "quality_length":
        data = data[:len(data) // 2] + data[len(data) // 2:].replace(b"IIII\n", b"III\n", 1)
    else:
        data = data[:len(data) // 2] + data[len(data) // 2:].replace(b"\n+\n", b"\n+other\n", 1)
    p = str(tmp_path / "in.fastq")
    open(p, "wb").write(data)
    r = subprocess.run([tool, "stream", p, "-", "4", str(1 << 20)], capture_output=True, text=True)
    assert r.returncode == 5, (what, r.returncode, r.stderr)


@pytest.mark.parametrize("variant", ["plain", "crlf", "blank_lines", "plus_header", "gzip", "cr_only"])
def test_reader_fastq_variants(tool, tmp_path, variant):
    """in_reads.cpp:181-221: '\\n' and '\\r' both end a line, empty lines are skipped, '+' line empty or equal to the header."""
    data = _fastq(RECORDS, eol={"crlf": "\r\n", "cr_only": "\r"}.get(variant, "\n"), plus_header=(0, 2) if variant == "plus_header" else ())
    if variant == "blank_lines":
        data = data.replace(b"\n@r2", b"\n\n\n@r2") + b"\n\n"
    p = str(tmp_path / ("in.fastq.gz" if variant == "gzip" else "in.fastq"))
    if variant == "gzip":
        with gzip.open(p, "wb") as f:
            f.write(data)
    else:
        open(p, "wb").write(data)
    r, st = _parse(tool, p, str(tmp_path / "o"))
    assert r.returncode == 0, r.stderr
    assert st["is_fastq"] == 1 and st["n_reads"] == 3 and st["is_gzip"] == (1 if variant == "gzip" else 0)
    assert st["total_bytes"] == len(data) and st["total_bases"] == 20
    # header statistics count the '@' / '+' lines as the reference does (in_reads.cpp:49, :81)
    assert st["total_symb_header"] == sum(len(h) + 1 for h, _, _ in RECORDS) + 3 + (sum(len(RECORDS[i][0]) for i in (0, 2)) if variant == "plus_header" else 0)
    assert open(str(tmp_path / "o.bases"), "rb").read() == b"".join(s.encode() for _, s, _ in RECORDS)
    assert open(str(tmp_path / "o.quals"), "rb").read() == b"".join(q.encode() for _, _, q in RECORDS)
    assert open(str(tmp_path / "o.headers"), "rb").read() == b"".join(h.encode() for h, _, _ in RECORDS)
    assert list(np.fromfile(str(tmp_path / "o.offsets"), np.uint64)) == [0, 9, 13, 20]
    assert list(np.fromfile(str(tmp_path / "o.plus"), np.uint8)) == ([1, 0, 1] if variant == "plus_header" else [0, 0, 0])
    assert list(np.fromfile(str(tmp_path / "o.hasn"), np.uint8)) == [1, 0, 0]
    assert st["read_packs"] == [3] and st["header_packs"] == [3]


@pytest.mark.parametrize("variant,message", [
    ("lowercase", "Only ACGTN symbols supported inside a read"), ("iupac", "Only ACGTN symbols supported inside a read"),
    ("bad_plus", "quality header not empty but different than read header"), ("no_final_eol", "something went wrong during input reading"),
    ("empty", "is empty"), ("unknown", "unknown file format"), ("truncated_record", "something went wrong during input reading")])
def test_reader_refusals(tool, tmp_path, variant, message):
    """The inputs the reference exits on (in_reads.cpp:31-35, :86-92, :246-260, :278-282) are refused with its messages."""
    data = {
        "lowercase": _fastq([("r", "ACgT", "IIII")]), "iupac": _fastq([("r", "ACRT", "IIII")]),
        "bad_plus": b"@r1\nACGT\n+r2\nIIII\n", "no_final_eol": _fastq(RECORDS)[:-1], "empty": b"", "unknown": b"ACGT\n",
        "truncated_record": b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n",
    }[variant]
    p = str(tmp_path / "in.fastq")
    open(p, "wb").write(data)
    r, _ = _parse(tool, p, str(tmp_path / "o"))
    assert r.returncode == 1 and message in r.stderr, (r.returncode, r.stderr)


def test_reader_fasta_multiline(tool, tmp_path):
    """in_reads.cpp:117-176: reads may span lines; '>' starts a header only at the start of a line; the last read needs no end-of-line."""
    data = b">s1 desc\nACGT\nACGT\r\n\r\nAC\n>s2\nNNNN\n>s3\nGG\nTT"
    p = str(tmp_path / "in.fa")
    open(p, "wb").write(data)
    r, st = _parse(tool, p, str(tmp_path / "o"))
    assert r.returncode == 0, r.stderr
    assert st["is_fastq"] == 0 and st["n_reads"] == 3 and st["total_bases"] == 18 and st["total_symb_header"] == 8 + 3 + 3
    assert open(str(tmp_path / "o.bases"), "rb").read() == b"ACGTACGTACNNNNGGTT"
    assert open(str(tmp_path / "o.headers"), "rb").read() == b"s1 descs2s3"
    assert open(str(tmp_path / "o.quals"), "rb").read() == b""
    assert list(np.fromfile(str(tmp_path / "o.hasn"), np.uint8)) == [0, 1, 0]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="the stock reference binary is not built here")
@pytest.mark.parametrize("lens", [[50, 900, 300, 1200, 70], [900, 50, 1200], [40, 40, 40], [500]])
def test_statistics_block_on_reads_without_overlaps(tool, tmp_path, lens):
    """stats_report.h against the stock binary's -v block on inputs whose reads share nothing (every read comes out plain, reads with N
    counted apart): all lines equal — including the reference's rule that a read is only tried as the minimum when it is not a new
    maximum (stats_collector.cpp:152-160: the first read never is, so `min read len` of [50, 900, ...] is not 50)."""
    rng = np.random.default_rng(len(lens))
    recs = []
    for i, n in enumerate(lens):
        b = rng.integers(0, 4, n)
        seq = bytearray(b"ACGT"[x] for x in b)
        if i == 2:
            seq[n // 2] = ord("N")
        recs.append(b"@r%d\n" % i + bytes(seq) + b"\n+\n" + b"5" * n + b"\n")
    inp = str(tmp_path / "in.fastq")
    open(inp, "wb").write(b"".join(recs))
    rr = subprocess.run([REF_BIN, "compress-ont", "-v", "-t", "2", inp, str(tmp_path / "ref.colord")], capture_output=True, text=True, cwd=str(tmp_path))
    assert rr.returncode == 0, rr.stderr[-1000:]
    lines = rr.stderr.splitlines()
    want = [l.rstrip() for l in lines[next(i for i, l in enumerate(lines) if "READS STATS" in l):] if l.strip()]
    got = [l.rstrip() for l in subprocess.run([tool, "stats", inp], capture_output=True, text=True, check=True).stdout.splitlines() if l.strip()]
    assert got == want


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="the stock reference binary is not built here")
def test_statistics_block_format_with_levels(tool, tmp_path):
    """The text of the whole block, levels included: the stock binary's block on overlapping reads, its numbers fed to the host printer
    (stats_report.h), gives the same text back."""
    s = synth.generate(400, 40000, 2500, seed=31, profile="ont", n_frac=0.03)
    inp = str(tmp_path / "in.fastq")
    s.write_fastq(inp)
    rr = subprocess.run([REF_BIN, "compress-ont", "-v", "-t", "2", inp, str(tmp_path / "ref.colord")], capture_output=True, text=True, cwd=str(tmp_path))
    assert rr.returncode == 0, rr.stderr[-1000:]
    lines = rr.stderr.splitlines()
    want = [l for l in lines[next(i for i, l in enumerate(lines) if "READS STATS" in l):] if l.strip()]
    assert sum("level" in l for l in want) >= 2
    nums = str(tmp_path / "nums.txt")
    open(nums, "w").write(" ".join(l.split(" : ")[1] for l in want if " : " in l))
    got = [l for l in subprocess.run([tool, "stats-format", nums], capture_output=True, text=True, check=True).stdout.splitlines() if l.strip()]
    assert got == want


def test_reader_packs_follow_the_reference_rule(tool, tmp_path):
    """Read packs close at >= 4 MiB of reads (one guard byte each), header packs at >= 4 MiB of headers."""
    rng = np.random.default_rng(1)
    lens = rng.integers(150000, 250000, 60)
    recs = [("read%d" % i, "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(n))), "I" * int(n)) for i, n in enumerate(lens)]
    p = str(tmp_path / "in.fastq")
    open(p, "wb").write(_fastq(recs))
    r, st = _parse(tool, p, str(tmp_path / "o"))
    assert r.returncode == 0, r.stderr
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    assert st["read_packs"] == _packs(off) and len(st["read_packs"]) > 2 and sum(st["read_packs"]) == 60
    assert st["header_packs"] == [60]


# ---------------------------------------------------------------------------------------- device-made archives (fixtures)
B200 = os.path.join(ROOT, "tests", "golden", "b200_archives")


@pytest.mark.skipif(not os.path.exists(os.path.join(B200, "expected.json")), reason="device-made archive fixtures not generated yet")
@pytest.mark.parametrize("case", sorted(json.load(open(os.path.join(B200, "expected.json")))) if os.path.exists(os.path.join(B200, "expected.json")) else [])
def test_decompress_device_made_archives(tmp_path, case):
    """colord-b200 decompress (host code, runs without a GPU) of archives the device path wrote on a B200: the FASTQ comes back
    with bases and headers identical and the qualities the reference's decompressor prints for that mode."""
    import hashlib
    cli = os.path.join(ROOT, "colord_b200", "colord-b200")
    if not os.path.exists(cli):
        subprocess.run(["make", "-C", os.path.join(ROOT, "colord_b200", "csrc")], check=True, capture_output=True)
    exp = json.load(open(os.path.join(B200, "expected.json")))[case]
    out = str(tmp_path / "back")
    r = subprocess.run([cli, "decompress", os.path.join(B200, case + ".colord"), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    data = open(out, "rb").read()
    assert len(data) == exp["output_bytes"] and hashlib.sha1(data).hexdigest() == exp["output_sha1"]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "colord")) or not os.path.exists(os.path.join(B200, "expected.json")),
                    reason="needs the reference binary (oracle/_ref, build container only) and the device-made fixtures")
def test_reference_binary_on_device_made_archive(tmp_path):
    """The unmodified reference reads the container and the info record of a device-made archive (`colord info`), and refuses to
    decompress it at its version check instead of misreading the native streams."""
    ref = os.path.join(ROOT, "oracle", "_ref", "colord")
    a = os.path.join(B200, "ont_default.colord")
    r = subprocess.run([ref, "info", a], capture_output=True, text=True)
    text = r.stdout + r.stderr
    assert r.returncode == 0 and "version major: 201" in text and "total reads: 40" in text and "total bases: 199305" in text and "compress-ont" in text
    r = subprocess.run([ref, "decompress", a, str(tmp_path / "x.fastq")], capture_output=True, text=True)
    assert r.returncode == 1 and "incompatibile archive version" in (r.stdout + r.stderr)


@pytest.mark.skipif(not os.path.exists(os.path.join(B200, "expected.json")), reason="device-made archive fixtures not generated yet")
def test_api_surface_on_device_made_archives(tmp_path):
    """include/colord_b200_api.h (mirror of src/API/colord_api.h): the reference's API example, ported by renaming the namespace,
    prints the records of every fixture archive exactly as `decompress` writes them, and the Info block of the reference's format."""
    import hashlib
    exe = str(tmp_path / "api_example")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "host_api_example.cpp"), "-lz"], check=True)
    exp = json.load(open(os.path.join(B200, "expected.json")))
    for case, e in exp.items():
        r = subprocess.run([exe, os.path.join(B200, case + ".colord")], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert hashlib.sha1(r.stdout).hexdigest() == e["output_sha1"], case
        info = r.stderr.decode()
        assert "colord archive version: 201.1.0" in info and f"is fastq: {'false' if case == 'ont_fasta' else 'true'}" in info
        assert ("reads source: PBRaw" if case.startswith("clr") else "reads source: ONT") in info
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "archives", "ont_default.colord")], capture_output=True, text=True)
    # an archive of the unmodified reference (6 reads of its ONT test file) is read through the same surface (compat_decoder.h)
    assert r.returncode == 0 and "colord archive version: 1.2.1" in r.stderr and r.stdout.count("\n") == 24, r.stderr
    # ... also one that carries its reference genome; one that does not asks for the genome (the API's second constructor)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "archives", "ont_genome_stored.colord")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("\n") == 24, r.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "archives", "ont_genome_checksum.colord")], capture_output=True, text=True)
    assert r.returncode == 1 and "reference genome is required" in r.stderr


@pytest.mark.parametrize("variant", ["plain", "crlf", "blank_lines", "plus_header", "tricky_quals", "gzip"])
def test_reader_threaded_equals_serial(tool, tmp_path, variant):
    """The threaded FASTQ parse (pieces cut at record starts) gives exactly the serial parser's arrays, statistics and packs —
    also when quality lines begin with '@' or '+', the case the record-start rule exists for."""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(400):
        n = int(rng.integers(1, 300))
        seq = "".join("ACGTN"[int(x)] for x in rng.choice(5, n, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        q = "".join(chr(int(x)) for x in rng.integers(33, 127, n))
        if variant == "tricky_quals":
            q = ("@" if i % 3 == 0 else "+" if i % 3 == 1 else "I") + q[1:]
        recs.append(("read%d/%d x=@+" % (i, n), seq, q))
    data = _fastq(recs, eol="\r\n" if variant == "crlf" else "\n", plus_header=range(0, 400, 7) if variant == "plus_header" else ())
    if variant == "blank_lines":
        data = data.replace(b"\n@read200", b"\n\n\n@read200").replace(b"\n+\n", b"\n\n+\n", 50)
    p = str(tmp_path / ("in.fastq.gz" if variant == "gzip" else "in.fastq"))
    if variant == "gzip":
        with gzip.open(p, "wb") as f:
            f.write(data)
    else:
        open(p, "wb").write(data)
    outs = {}
    for name, threads in (("serial", 1), ("threaded", 7)):
        r = subprocess.run([tool, "parse", p, str(tmp_path / name), str(threads), "2000"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[name] = json.loads(r.stdout)
    assert outs["serial"]["threads_used"] == 1 and outs["threaded"]["threads_used"] >= 4
    for key in ("n_reads", "total_bytes", "total_bases", "total_symb_header", "read_packs", "header_packs"):
        assert outs["serial"][key] == outs["threaded"][key], key
    assert outs["serial"]["n_reads"] == 400
    for ext in ("bases", "offsets", "quals", "headers", "hoff", "plus", "hasn"):
        assert open(str(tmp_path / ("serial." + ext)), "rb").read() == open(str(tmp_path / ("threaded." + ext)), "rb").read(), ext
    assert open(str(tmp_path / "serial.bases"), "rb").read() == "".join(s for _, s, _ in recs).encode()
    assert open(str(tmp_path / "serial.quals"), "rb").read() == "".join(q for _, _, q in recs).encode()


@pytest.mark.parametrize("variant,message", [("bad_symbol", "Only ACGTN symbols supported inside a read"), ("bad_plus", "quality header not empty but different than read header"),
                                             ("no_final_eol", "something went wrong during input reading"), ("missing_line", "Only ACGTN symbols supported inside a read")])
def test_reader_threaded_refusals_are_the_serial_ones(tool, tmp_path, variant, message):
    recs = [("r%d" % i, "ACGT" * 20, "I" * 80) for i in range(300)]
    data = _fastq(recs)
    if variant == "bad_symbol":
        data = data.replace(b"@r250\nACGT", b"@r250\nACgT")
    elif variant == "bad_plus":
        data = data.replace(b"@r250\n" + b"ACGT" * 20 + b"\n+\n", b"@r250\n" + b"ACGT" * 20 + b"\n+r251\n")
    elif variant == "no_final_eol":
        data = data[:-1]
    else:
        data = data.replace(b"@r250\n", b"", 1)
    p = str(tmp_path / "in.fastq")
    open(p, "wb").write(data)
    r = subprocess.run([tool, "parse", p, str(tmp_path / "o"), "6", "2000"], capture_output=True, text=True)
    assert r.returncode == 1 and message in r.stderr, (r.returncode, r.stderr)
    r1 = subprocess.run([tool, "parse", p, str(tmp_path / "o1"), "1", "2000"], capture_output=True, text=True)
    assert (r1.returncode, r1.stderr) == (r.returncode, r.stderr)          # a shifted record reads the next line as the read, as in the reference


def test_cli_without_a_gpu_fails_loudly(tmp_path):
    """No CPU fallback behind the command line either: without a CUDA device compress-* stops with the library's message and
    exit code 1 (on a GPU box the same command compresses: tests/test_gpu_cli.py); decompress / info need no device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cli = os.path.join(ROOT, "colord_b200", "colord-b200")
    if not os.path.exists(cli):
        subprocess.run(["make", "-C", os.path.join(ROOT, "colord_b200", "csrc")], check=True, capture_output=True)
    p = str(tmp_path / "in.fastq")
    open(p, "wb").write(b"@r1\nACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIII\n")
    r = subprocess.run([cli, "compress-ont", p, str(tmp_path / "o.colord")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr
    r = subprocess.run([cli, "info", os.path.join(ROOT, "tests", "golden", "archives", "ont_default.colord")], capture_output=True, text=True)
    assert r.returncode == 0 and "total reads: 6" in r.stderr


READER_CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "reader_cases.json")))


@pytest.mark.parametrize("case", sorted(READER_CASES))
def test_reader_against_the_reference_binary(tool, tmp_path, case):
    """Differential test against the UNMODIFIED reference (fixtures by make_reader_golden.py): for every input the reference's
    `compress-ont -q org` accepts, the reader accepts it too and holds exactly the records the reference's `decompress` writes
    back; every input the reference refuses is refused, with the reference's message where that message is its reader's.
    Inputs on which the reference segfaults or aborts on an assertion (truncated records, some blank-line layouts) carry no
    verdict: there the reader only has to end in a clean accept or refusal."""
    import base64
    c = READER_CASES[case]
    data = base64.b64decode(c["input_b64"])
    p = str(tmp_path / ("in.fastq.gz" if case == "gzip" else "in.fa" if data[:1] == b">" else "in.fastq"))
    open(p, "wb").write(data)
    r, st = _parse(tool, p, str(tmp_path / "o"))
    if c.get("crashed"):
        assert r.returncode in (0, 1)
        return
    if not c["accepted"]:
        assert r.returncode == 1, (case, r.stderr)
        for phrase in ("Only ACGTN symbols supported inside a read", "something went wrong during input reading", "is empty", "unknown file format"):
            if phrase in c["error"]:
                assert phrase in r.stderr, (case, r.stderr)
        return
    assert r.returncode == 0, (case, r.stderr)
    bases = open(str(tmp_path / "o.bases"), "rb").read(); quals = open(str(tmp_path / "o.quals"), "rb").read(); hdr = open(str(tmp_path / "o.headers"), "rb").read()
    off = np.fromfile(str(tmp_path / "o.offsets"), np.uint64); hoff = np.fromfile(str(tmp_path / "o.hoff"), np.uint64); plus = np.fromfile(str(tmp_path / "o.plus"), np.uint8)
    out = []
    for i in range(st["n_reads"]):
        h, s = hdr[int(hoff[i]):int(hoff[i + 1])], bases[int(off[i]):int(off[i + 1])]
        if st["is_fastq"]:
            out.append(b"@" + h + b"\n" + s + b"\n+" + (h if plus[i] else b"") + b"\n" + quals[int(off[i]):int(off[i + 1])] + b"\n")
        else:
            out.append(b">" + h + b"\n" + s + b"\n")
    assert b"".join(out) == base64.b64decode(c["output_b64"]), case


@pytest.mark.skipif(not os.path.exists(os.path.join(B200, "expected.json")), reason="device-made archive fixtures not generated yet")
def test_damaged_archives_never_crash_the_decompressor(tmp_path):
    """Bytes of device-made archives are overwritten at random places: `decompress` must end — with exit code 1 and a message, or
    (when the damage hits nothing that is checked) with a file — never with a signal or a hang."""
    cli = os.path.join(ROOT, "colord_b200", "colord-b200")
    if not os.path.exists(cli):
        subprocess.run(["make", "-C", os.path.join(ROOT, "colord_b200", "csrc")], check=True, capture_output=True)
    rng = np.random.default_rng(11)
    refused = 0
    for case in ("ont_default", "ont_org_small", "clr_ratio_none"):
        data = bytearray(open(os.path.join(B200, case + ".colord"), "rb").read())
        for trial in range(12):
            bad = bytearray(data)
            n_hits = int(rng.integers(1, 6))
            for _ in range(n_hits):
                at = int(rng.integers(0, len(bad)))
                span = int(rng.integers(1, 9))
                bad[at:at + span] = rng.integers(0, 256, min(span, len(bad) - at), dtype=np.uint8).tobytes()
            if trial % 4 == 3:
                bad = bad[:int(rng.integers(8, len(bad)))]          # truncation
            p = str(tmp_path / "bad.colord")
            open(p, "wb").write(bytes(bad))
            r = subprocess.run([cli, "decompress", p, str(tmp_path / "out")], capture_output=True, text=True, timeout=60)
            assert r.returncode in (0, 1), (case, trial, r.returncode, r.stderr[-200:])
            refused += r.returncode == 1
            if r.returncode == 1:
                assert r.stderr.strip(), (case, trial)
    assert refused >= 18          # most damage is noticed


@pytest.mark.skipif(not os.path.exists(os.path.join(B200, "expected.json")), reason="device-made archive fixtures not generated yet")
@pytest.mark.parametrize("mode,header", [(2, b"@@"), (1, b"@")])
def test_header_modes_none_and_main(tmp_path, mode, header):
    """-i none / -i main: no header bytes are stored and every record comes back as "@@" / "@" with an empty '+' line — what the
    unmodified reference writes for these modes (its `main` coder is an empty function).  The archive is made on the CPU from a
    device-made fixture by emptying the header stream and changing meta.headerComprMode (tests/host_archive_tool.cpp)."""
    tool2 = str(tmp_path / "atool")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", tool2, os.path.join(ROOT, "tests", "host_archive_tool.cpp")], check=True)
    cli = os.path.join(ROOT, "colord_b200", "colord-b200")
    a = str(tmp_path / "a.colord")
    subprocess.run([tool2, "set-header-mode", os.path.join(B200, "ont_org_small.colord"), a, str(mode)], check=True)
    full, got = str(tmp_path / "full"), str(tmp_path / "got")
    subprocess.run([cli, "decompress", os.path.join(B200, "ont_org_small.colord"), full], check=True)
    subprocess.run([cli, "decompress", a, got], check=True)
    want = open(full, "rb").read().split(b"\n")
    for i in range(0, len(want) - 1, 4):
        want[i] = header
        want[i + 2] = b"+"
    assert open(got, "rb").read() == b"\n".join(want)


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("case", ["ont_mem", "ont_bal", "clr_ratio", "hifi"])
def test_reference_tuples_through_the_dna_container_on_cpu(golden, tool, tmp_path, case, ragged):
    """The CompactES tuples the UNMODIFIED reference emitted (tests/golden/<case>/es.bin) -> the oracle's twin of the device's DNA
    encoder -> container "DB01" -> the two independent decoders (oracle/stage3_dna.c, host/decompressor.h) -> the input reads.
    Closes the loop on the CPU at levels 1, 2 and 3, sparse and all-reference modes, ONT / CLR / HiFi."""
    g = golden(case)
    s = g.reads_in
    packs = g.es_packs if sum(g.es_packs) == s.n_reads else [s.n_reads]
    if ragged:                                                            # packs of 1, 63, 64, 65 reads and the rest: lanes with 0, 1 and 2 reads
        packs = [1, 63, 64, 65, s.n_reads - 193]
    stream = oracle_lib.dna_encode(int(g.params["level"]), int(g.params["max_candidates"]), g.es, s.bases, s.offsets, g.is_ref, packs)
    assert len(stream) < s.n_bases                                        # it is a compressor
    bases, off = oracle_lib.dna_decode(stream, s.n_reads, g.is_ref, s.n_bases)
    assert np.array_equal(off, s.offsets) and np.array_equal(bases, s.bases)
    # the product decoder: decisions = the sampler's answer; reads holding N are dropped from the references inside
    dec = (g.is_ref | g.has_n).astype(np.uint8)
    paths = [_write(str(tmp_path / "s"), stream), str(s.n_reads), _write(str(tmp_path / "d"), dec)]
    ob, oo, of = str(tmp_path / "ob"), str(tmp_path / "oo"), str(tmp_path / "of")
    subprocess.run([tool, "dna", *paths, ob, oo, of], check=True)
    assert np.array_equal(np.fromfile(ob, np.uint8), s.bases) and np.array_equal(np.fromfile(oo, np.uint64), s.offsets)
    assert os.path.getsize(of) == s.n_bases
