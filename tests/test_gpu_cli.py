"""End to end through the command-line front end on a B200: FASTQ / FASTA file -> colord-b200 compress-* (host reader, C-ABI,
device stages 1-3, reference archive container) -> colord-b200 decompress (host decoders) -> file.
Inputs are rebuilt from committed fixtures (the reference's own test reads: tests/golden/qual_*.bin.gz + headers.json.gz) and
from the synthetic generator.  Expected output: bases and headers identical; qualities identical for `-q org`, equal to the
reference's .quan fixture for the default lossy modes, the constant the reference prints for `-q none`.
CLB_SAVE_ARCHIVES=<dir>: also keep the archives + the sha1 of the expected output (fixtures of tests/test_host_decode.py)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import golden_io
from colord_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colord_b200", "colord-b200")


@pytest.fixture(scope="module")
def cli():
    lib = os.path.join(ROOT, "colord_b200", "libcolord_b200.so")
    assert os.path.exists(lib), "libcolord_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    if not os.path.exists(CLI) or os.path.getmtime(CLI) < os.path.getmtime(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "colord_b200", "csrc"), "../colord-b200"], check=True, capture_output=True)
    return CLI


def _records(name, n):
    """(headers, bases, quals, expected lossy quals) of the first n reads of a reference test file, from the fixtures"""
    bases, quals, quan, off = golden_io.load_qual_golden(name)
    headers, _ = golden_io.load_hdr_golden()[name]
    n = min(n, len(off) - 1)
    cut = lambda a, i: a[int(off[i]):int(off[i + 1])].tobytes()
    return [(headers[i].lstrip(b"@"), cut(bases, i), cut(quals, i), cut(quan, i)) for i in range(n)]


def _synthetic(n, profile, seed):
    s = synth.generate(n, 60000, 3000, seed=seed, profile=profile, n_frac=0.05)
    cut = lambda a, i: a[int(s.offsets[i]):int(s.offsets[i + 1])].tobytes()
    return [(b"m64011_190830/%d/0_%d extra" % (4000 + 3 * i, int(s.offsets[i + 1] - s.offsets[i])), cut(s.bases, i), cut(s.quals, i), None) for i in range(s.n_reads)]


def _fastq(recs, qual_index, plus=False):
    return b"".join(b"@" + h + b"\n" + b + b"\n+" + (h if plus else b"") + b"\n" + r[qual_index] + b"\n" for r in recs for h, b in [(r[0], r[1])])


CASES = {
    # name: (records, command, options, which qualities come back, fasta?, '+' line repeats the header?)
    "ont_default": (lambda: _records("ont", 40), "compress-ont", [], "quan", False, False),
    "ont_org_balanced": (lambda: _records("ont", 100), "compress-ont", ["-q", "org", "-p", "balanced"], "org", False, True),
    "hifi_default": (lambda: _records("hifi", 40), "compress-pbhifi", [], "quan", False, False),
    "hifi_org_ratio": (lambda: _records("hifi", 12), "compress-pbhifi", ["-q", "org", "-p", "ratio"], "org", False, False),
    "ont_org_small": (lambda: _records("ont", 20), "compress-ont", ["-q", "org", "-p", "balanced"], "org", False, True),
    "clr_ratio_none": (lambda: _synthetic(120, "clr", 4), "compress-pbraw", ["-p", "ratio"], "none", False, False),
    "ont_fasta": (lambda: _records("ont", 60), "compress-ont", ["-p", "balanced"], None, True, False),
    "ont_2avg_k18": (lambda: _synthetic(150, "ont", 9), "compress-ont", ["-q", "2-avg", "-k", "18", "-a", "15", "-R", "all"], "2avg", False, False),
    # -i none / main: the reference stores no header bytes and prints "@@" / "@" (tests/test_host_decode.py::test_header_modes_none_and_main)
    "ont_hdr_none": (lambda: _records("ont", 30), "compress-ont", ["-i", "none", "-q", "org"], "org", False, False),
    "ont_hdr_main": (lambda: _records("ont", 30), "compress-ont", ["-i", "main", "-q", "org"], "org", False, False),
}


REF = os.path.join(ROOT, "oracle", "_ref", "colord")      # the unmodified reference (oracle/Makefile); travels to the GPU box


def _stream_sizes(stderr):
    import re
    return {k: int(v) for k, v in re.findall(r"^(DNA|Quality|Header) size\s*:\s*(\d+)", stderr, re.M)}


def _same_parts(ours, theirs, what):
    """the three streams of two archives hold the same parts: metadata and payload bytes (the printed sizes also count the footer's
    varints, which move with the order the streams' parts were appended in)"""
    import colord_archive
    a, b = colord_archive.read_parts(ours), colord_archive.read_parts(theirs)
    for stream in ("dna", "qual", "header"):
        assert (stream in a) == (stream in b), (what, stream)
        if stream in a:
            assert [md for md, _ in a[stream]] == [md for md, _ in b[stream]], (what, stream, "part metadata")
            assert [len(x) for _, x in a[stream]] == [len(x) for _, x in b[stream]], (what, stream, "part sizes")
            assert all(x == y for (_, x), (_, y) in zip(a[stream], b[stream])), (what, stream, "part bytes")


@pytest.mark.parametrize("fmt", ["native", "compat"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_cli_round_trip(cli, tmp_path, case, fmt):
    """fmt native: the device's containers; compat: the reference's own streams — the archive is then also decompressed by the
    unmodified reference, and its three streams have exactly the sizes the reference's own compression of the file gives."""
    make, cmd, opts, qmode, fasta, plus = CASES[case]
    opts = opts + ["--" + fmt]
    recs = make()
    inp = str(tmp_path / ("in.fa" if fasta else "in.fastq"))
    if fasta:
        data = b"".join(b">" + h + b"\n" + b + b"\n" for h, b, _, _ in recs)
        want = data
    else:
        data = _fastq(recs, 2, plus)
        if qmode == "org":
            want = data
        elif qmode == "quan":
            want = _fastq(recs, 3, plus)
        elif qmode == "none":
            want = _fastq([(h, b, b"!" * len(b), None) for h, b, _, _ in recs], 2, plus)
        else:       # 2-avg has no reference fixture: the oracle's restatement of the reference's transform is the expectation
            import oracle_lib
            P = oracle_lib.qual_params(2, [7], 1)
            bases = np.frombuffer(b"".join(r[1] for r in recs), np.uint8)
            quals = np.frombuffer(b"".join(r[2] for r in recs), np.uint8)
            off = np.concatenate([[0], np.cumsum([len(r[1]) for r in recs])]).astype(np.uint64)
            lossy = oracle_lib.qual_lossy(P, bases, quals, off)
            want = _fastq([(r[0], r[1], lossy[int(off[i]):int(off[i + 1])].tobytes(), None) for i, r in enumerate(recs)], 2, plus)
    if "-i" in opts:
        hdr = b"@" if opts[opts.index("-i") + 1] == "none" else b""
        want = _fastq([(hdr, r[1], r[2], None) for r in recs], 2, False)
    open(inp, "wb").write(data)
    arch, back = str(tmp_path / "a.colord"), str(tmp_path / "back")
    r = subprocess.run([cli, cmd, *opts, inp, arch], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "DNA size" in r.stderr and "Header size" in r.stderr
    ours = _stream_sizes(r.stderr)
    if fmt == "compat" and os.path.exists(REF):
        ref_arch, ref_back = str(tmp_path / "ref.colord"), str(tmp_path / "ref_back")
        rr = subprocess.run([REF, cmd, *opts[:-1], "-t", "4", inp, ref_arch], capture_output=True, text=True, cwd=str(tmp_path))
        assert rr.returncode == 0, rr.stderr
        _same_parts(arch, ref_arch, case)
        rr = subprocess.run([REF, "decompress", arch, ref_back], capture_output=True, text=True, cwd=str(tmp_path))
        assert rr.returncode == 0, rr.stderr
        assert open(ref_back, "rb").read() == want, f"{case}: the reference's decompress of our archive differs"
    r = subprocess.run([cli, "decompress", arch, back], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = open(back, "rb").read()
    assert got == want, f"{case}: round trip differs ({len(got)} vs {len(want)} bytes)"
    if fmt == "compat" or qmode != "org" or len(recs) >= 100:       # a few reads of lossless qualities do not pay for their 96-symbol tables
        assert os.path.getsize(arch) < len(data)
    # the info record as the reference's `colord info` prints it
    r = subprocess.run([cli, "info", arch], capture_output=True, text=True)
    assert r.returncode == 0 and f"total reads: {len(recs)}" in r.stderr and f"total bases: {sum(len(x[1]) for x in recs)}" in r.stderr and f"total bytes: {len(data)}" in r.stderr
    save = os.environ.get("CLB_SAVE_ARCHIVES")
    if save and fmt == "native" and case in ("ont_default", "ont_org_small", "ont_fasta", "clr_ratio_none"):
        os.makedirs(save, exist_ok=True)
        open(os.path.join(save, case + ".colord"), "wb").write(open(arch, "rb").read())
        p = os.path.join(save, "expected.json")
        exp = json.load(open(p)) if os.path.exists(p) else {}
        exp[case] = {"made_by": " ".join(["colord-b200", cmd, *opts]), "output_bytes": len(want), "output_sha1": hashlib.sha1(want).hexdigest(), "archive_bytes": os.path.getsize(arch)}
        json.dump(exp, open(p, "w"), indent=1, sort_keys=True)


@pytest.mark.skipif(not os.path.exists(REF), reason="the stock reference binary is not built here")
@pytest.mark.parametrize("opts", [["-q", "2-fix"], ["-q", "4-fix", "-p", "balanced"], ["-q", "5-fix", "-T", "6,13,25,50"], ["-q", "avg"], ["-q", "5-avg", "-p", "ratio"], ["-q", "none", "-c", "9"]])
def test_cli_compat_every_quality_mode_against_the_reference(cli, tmp_path, opts):
    """The quality modes only the compat streams hold (threshold modes, plain average) and a few more option sets: stream sizes equal
    the reference's, and the four ways through (ours / the reference's compressor x ours / the reference's decompressor) print one file."""
    s = synth.generate(900, 150000, 2500, seed=17, profile="ont", n_frac=0.03)
    inp = str(tmp_path / "in.fastq")
    s.write_fastq(inp)
    a, b = str(tmp_path / "ours.colord"), str(tmp_path / "ref.colord")
    r = subprocess.run([cli, "compress-ont", *opts, "--compat", inp, a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref_opts = [x for o in opts for x in (o.split(",") if "," in o else [o])]      # the reference takes -T as separate numbers
    rr = subprocess.run([REF, "compress-ont", *ref_opts, "-t", "4", inp, b], capture_output=True, text=True, cwd=str(tmp_path))
    assert rr.returncode == 0, rr.stderr
    _same_parts(a, b, "-".join(opts))
    outs = []
    for exe, arc in ((cli, a), (cli, b), (REF, a), (REF, b)):
        o = str(tmp_path / f"out{len(outs)}")
        assert subprocess.run([exe, "decompress", arc, o], capture_output=True, cwd=str(tmp_path)).returncode == 0
        outs.append(open(o, "rb").read())
    assert outs[0] == outs[1] == outs[2] == outs[3]


def _stats_block(stderr):
    """the statistics block of a verbose compression: from the READS STATS banner to its last counter line"""
    lines = stderr.splitlines()
    start = next(i for i, l in enumerate(lines) if "READS STATS" in l)
    return [l.rstrip() for l in lines[start:] if l.startswith(" * * *") or l.startswith(" ----") or (l.startswith(("#", "min read", "max read")) and " : " in l)]


@pytest.mark.skipif(not os.path.exists(REF), reason="the stock reference binary is not built here")
@pytest.mark.parametrize("mode,opts,profile", [("compress-ont", [], "ont"), ("compress-ont", ["-p", "ratio", "-c", "8"], "ont"), ("compress-pbhifi", [], "hifi"), ("compress-pbraw", ["-R", "all"], "clr")])
def test_cli_verbose_statistics_equal_the_reference(cli, tmp_path, mode, opts, profile):
    """-v: the statistics block the reference prints when it exits (stats_collector.cpp:100-146: reads, refuse reasons, plain reads,
    orientation choices, and per recursion level alternatives / plain symbols / edit-script symbol classes / anchors / flanks) — every
    line equals the stock binary's on the same file, in both stream formats."""
    s = synth.generate(1500, 120000, 3000, seed=23, profile=profile, n_frac=0.04)
    inp = str(tmp_path / "in.fastq")
    s.write_fastq(inp)
    rr = subprocess.run([REF, mode, *opts, "-v", "-t", "4", inp, str(tmp_path / "ref.colord")], capture_output=True, text=True, cwd=str(tmp_path))
    assert rr.returncode == 0, rr.stderr[-2000:]
    want = _stats_block(rr.stderr)
    assert any("level 0" in l for l in want) and len(want) > 30, want
    for fmt in ("--native", "--compat"):
        r = subprocess.run([cli, mode, *opts, fmt, "-v", inp, str(tmp_path / "ours.colord")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        got = _stats_block(r.stderr)
        assert got == want, "\n".join(f"{a!r} | {b!r}" for a, b in zip(got, want) if a != b)


def test_cli_streamed_input_equals_whole_file_input(cli, tmp_path):
    """Files of 64 MiB and more stream to the device in pieces while they are parsed (qualities resident on the device); the archive's
    streams must be the ones the whole-file path writes (CLB_NO_STREAMING=1), and the round trip lossless."""
    import colord_archive
    fq = str(tmp_path / "in.fastq")
    synth.generate_file(fq, "ont", 4500, 1_700_000, 8000, seed=8, workers=8)
    assert os.path.getsize(fq) >= 64 << 20
    arcs = {}
    for name, env in (("streamed", {}), ("whole", {"CLB_NO_STREAMING": "1"})):
        arcs[name] = str(tmp_path / (name + ".colord"))
        r = subprocess.run([cli, "compress-ont", "-q", "org", "-v", "--native", fq, arcs[name]], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        assert ("streamed to the device" in r.stderr) == (name == "streamed"), r.stderr
    a, b = colord_archive.read_parts(arcs["streamed"]), colord_archive.read_parts(arcs["whole"])
    for stream in ("dna-b200", "qual-b200", "header-b200", "meta"):
        assert a[stream] == b[stream], stream
    back = str(tmp_path / "back")
    assert subprocess.run([cli, "decompress", arcs["streamed"], back], capture_output=True).returncode == 0
    assert subprocess.run(["cmp", "-s", fq, back]).returncode == 0
    r = subprocess.run([cli, "compress-ont", "--compat", fq, str(tmp_path / "c.colord")], capture_output=True, text=True)      # compat streams from resident qualities
    assert r.returncode == 0, r.stderr
    if os.path.exists(REF):
        rr = subprocess.run([REF, "compress-ont", "-t", "8", fq, str(tmp_path / "r.colord")], capture_output=True, text=True, cwd=str(tmp_path))
        assert rr.returncode == 0, rr.stderr
        _same_parts(str(tmp_path / "c.colord"), str(tmp_path / "r.colord"), "streamed compat")


@pytest.mark.parametrize("fmt", ["native", "compat"])
@pytest.mark.parametrize("store", [True, False])
def test_cli_reference_genome_mode(cli, tmp_path, fmt, store):
    """-G [-s] (BASELINE config 5): the genome's sequences are counted with the reads, its pseudo-reads are the first reference reads.
    Round trip in both stream formats; in compat format the archive's parts (the stored genome included) are the stock binary's and
    each decompressor reads the other's archive."""
    import importlib.util
    import colord_archive
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_ref_archive_golden.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    gen = dict(n_reads=700, genome_len=200000, mean_len=2500, seed=41, profile="ont", n_frac=0.03)
    s = synth.generate(**gen)
    fq, genome = str(tmp_path / "in.fastq"), str(tmp_path / "genome.fa")
    s.write_fastq(fq)
    mk.write_genome(genome, gen)
    opts = ["-G", genome, "-q", "org"] + (["-s"] if store else [])
    dopts = [] if store else ["-G", genome]
    arch, back = str(tmp_path / "a.colord"), str(tmp_path / "back")
    r = subprocess.run([cli, "compress-ont", *opts, "--" + fmt, "-v", fq, arch], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "# ref genome pseudo reads:" in r.stderr
    r = subprocess.run([cli, "decompress", *dopts, arch, back], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(back, "rb").read() == open(fq, "rb").read()
    # the genome pays: the DNA stream is smaller than without it
    plain = str(tmp_path / "p.colord")
    assert subprocess.run([cli, "compress-ont", "-q", "org", "--" + fmt, fq, plain], capture_output=True).returncode == 0
    with_g, without = colord_archive.read_parts(arch), colord_archive.read_parts(plain)
    dna = "dna" if fmt == "compat" else "dna-b200"
    assert sum(len(b) for _, b in with_g[dna]) < sum(len(b) for _, b in without[dna])
    if fmt == "compat" and os.path.exists(REF):
        ref_arch, ref_back = str(tmp_path / "ref.colord"), str(tmp_path / "ref_back")
        rr = subprocess.run([REF, "compress-ont", *opts, "-t", "4", fq, ref_arch], capture_output=True, text=True, cwd=str(tmp_path))
        assert rr.returncode == 0, rr.stderr
        _same_parts(arch, ref_arch, "genome")
        if store:
            assert with_g["ref-genome"] == colord_archive.read_parts(ref_arch)["ref-genome"]
        rr = subprocess.run([REF, "decompress", *dopts, arch, ref_back], capture_output=True, text=True, cwd=str(tmp_path))
        assert rr.returncode == 0, rr.stderr
        assert open(ref_back, "rb").read() == open(fq, "rb").read()


def _n_gpus():
    try:
        r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
        return sum(1 for l in r.stdout.splitlines() if l.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs on the box (gpurun --gpus 2)")
@pytest.mark.parametrize("n_gpus", sorted({2, max(2, _n_gpus())}))
@pytest.mark.parametrize("opts", [[], ["-q", "org", "-p", "balanced"], ["-R", "all", "-q", "none"]])
def test_cli_multi_gpu_archive_round_trip(cli, tmp_path, opts, n_gpus):
    """--gpus 2: the input is sharded over two GPUs (NCCL exchanges of the k-mer counts and the reference reads in
    libcolord_b200_mgpu.so), the archive holds one part per stream and GPU, and `decompress` gives the file back.  The DNA stream
    must stay close to the single-GPU archive's (the shards see the same reference reads; only tables and pack cuts differ)."""
    import colord_archive
    fq = str(tmp_path / "in.fastq")
    synth.generate_file(fq, "ont", 3000, 1_100_000, 8000, seed=12, workers=8)
    one, two, back = str(tmp_path / "one.colord"), str(tmp_path / "two.colord"), str(tmp_path / "back")
    r = subprocess.run([cli, "compress-ont", *opts, "--native", "-v", fq, one], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stats_one = _stats_block(r.stderr)
    r = subprocess.run([cli, "compress-ont", *opts, "--gpus", str(n_gpus), "-v", fq, two], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # -v: the ranks' counters add up to the one-GPU report — exactly for the read statistics, refuse reasons, plain reads and orientation
    # choices (lines up to the first level); the per-level lines follow the estimator, which restarts at the shards' pack cuts
    stats_n = _stats_block(r.stderr)
    first_level = next(i for i, l in enumerate(stats_one) if "level 0" in l)
    assert stats_n[:first_level] == stats_one[:first_level] and len(stats_n) == len(stats_one) > 30, (stats_n[:first_level], stats_one[:first_level])
    for a, b in zip(stats_n[first_level:], stats_one[first_level:]):
        if " : " in a:
            va, vb = int(a.split(" : ")[1]), int(b.split(" : ")[1])
            assert a.split(" : ")[0] == b.split(" : ")[0] and abs(va - vb) <= 0.02 * max(va, vb) + 64, (a, b)
    parts1, parts2 = colord_archive.read_parts(one), colord_archive.read_parts(two)
    assert len(parts2["dna-b200"]) == n_gpus and len(parts1["dna-b200"]) == 1
    assert sum(md for md, _ in parts2["dna-b200"]) == parts1["dna-b200"][0][0]
    d1, d2 = len(parts1["dna-b200"][0][1]), sum(len(b) for _, b in parts2["dna-b200"])
    assert d2 < (1.03 if n_gpus == 2 else 1.20) * d1, (d1, d2)      # 24 Mbases over N shards: every shard pays its own tables
    r = subprocess.run([cli, "decompress", two, back], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if "org" in opts:
        assert subprocess.run(["cmp", "-s", fq, back]).returncode == 0
    else:      # lossy / no qualities: bases and headers must be the input's, the qualities the single-GPU archive's
        back1 = str(tmp_path / "back1")
        assert subprocess.run([cli, "decompress", one, back1], capture_output=True).returncode == 0
        assert subprocess.run(["cmp", "-s", back1, back]).returncode == 0


def test_cli_refusals(cli, tmp_path):
    """Errors end in exit code 1 with a message, as in the reference's CLI (arg_parse.cpp:820-902, in_reads.cpp)."""
    bad = str(tmp_path / "bad.fastq")
    open(bad, "wb").write(b"@r\nACXT\n+\nIIII\n")
    r = subprocess.run([cli, "compress-ont", bad, str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "Only ACGTN symbols supported inside a read" in r.stderr
    ok = str(tmp_path / "ok.fastq")
    open(ok, "wb").write(b"@r\nACGT\n+\nIIII\n")
    for opts, msg in ((["-G", "x.fa"], "cannot open file"), (["-q", "4-fix", "--native"], "compat streams only"), (["-q", "bogus"], "unknown quality"), (["-k", "99"], "15..28")):
        r = subprocess.run([cli, "compress-ont", *opts, ok, str(tmp_path / "o")], capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (opts, r.stderr)
    r = subprocess.run([cli, "decompress", ok, str(tmp_path / "o2")], capture_output=True, text=True)
    assert r.returncode == 1
