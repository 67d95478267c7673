#!/usr/bin/env python
"""tools/ratio_check.py — "matched ratio" and round trip of the command line against the STOCK reference binary, per BASELINE config.

For every configuration (BASELINE.json configs 1-4 + a north-star slice; the synthetic ones scaled by --scale) one FASTQ file is
written (colord_b200.synth.generate_file; C1 = the reference's bundled test/M.bovis.fastq rebuilt from the committed fixtures),
then on that same file:
  * `oracle/_ref/colord compress-* -t <nproc>`  (the unmodified reference, BASELINE.md §3: wall time process start -> exit; --runs N
    repeats and reports the median; one extra `-t 1` run with CPU seconds when --t1 is given)
  * `colord_b200/colord-b200 compress-*` with the same options
  * `colord-b200 decompress` of our archive, compared with the input: byte-identical for lossless configurations; for the lossy
    quality modes bases / headers identical and the qualities equal to the reference's lossy transform (oracle_lib.qual_lossy, pinned
    on the reference's own .quan files); C1 additionally against `colord decompress` of the reference's archive (cmp).
The verdict per configuration: archive bytes ours / reference <= 1.005 and the round trip green.  One JSON object per configuration
on stdout and in --out; a markdown table in --md.  Runs on the GPU box (needs a CUDA device for colord-b200); test infrastructure.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import resource
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.path.join(ROOT, "oracle", "_ref", "colord")
OURS = os.path.join(ROOT, "colord_b200", "colord-b200")

# name -> (mode, options, profile, reads at scale 1, mean read length, coverage, lossless?)       SURVEY.md §8d
CONFIGS = {
    "C1": ("compress-ont", ["-p", "memory", "-q", "4-avg"], None, 100, 0, 0, False),
    "C2": ("compress-pbhifi", ["-q", "org"], "hifi", 100_000, 15000, 15, True),
    "C3": ("compress-ont", ["-p", "balanced"], "ont", 1_000_000, 8000, 20, False),
    "C4": ("compress-pbraw", ["-q", "none", "-p", "ratio"], "clr", 2_000_000, 10000, 20, False),
    "NS": ("compress-ont", [], "ont", 3_125_000, 8000, 20.8, False),
}
SIZE_RE = re.compile(r"^(DNA|Quality|Header|Meta|Info) size\s*:\s*(\d+)", re.M)


def rebuild_m_bovis(path):
    import golden_io
    hdr = golden_io.load_hdr_golden()["ont"][0]
    bases, quals, _, off = golden_io.load_qual_golden("ont")
    with open(path, "wb") as f:
        for i, h in enumerate(hdr):
            a, b = int(off[i]), int(off[i + 1])
            f.write(h + b"\n" + bases[a:b].tobytes() + b"\n+\n" + quals[a:b].tobytes() + b"\n")
    return os.path.getsize(path), int(off[-1])


def run(cmd, **kw):
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)
    dt = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    cpu = (r1.ru_utime - r0.ru_utime) + (r1.ru_stime - r0.ru_stime)
    return p, dt, cpu


def sizes_of(stderr):
    return {k.lower(): int(v) for k, v in SIZE_RE.findall(stderr)}


def parse_fastq_arrays(path):
    """-> (headers bytes list, bases u8, quals u8, offsets u64); 4-line records"""
    import numpy as np
    raw = np.fromfile(path, np.uint8)
    nl = np.flatnonzero(raw == 10)
    assert len(nl) % 4 == 0, "not a 4-line FASTQ"
    starts = np.concatenate([[0], nl[:-1] + 1])
    h0, h1 = starts[0::4], nl[0::4]
    s0, s1 = starts[1::4], nl[1::4]
    q0, q1 = starts[3::4], nl[3::4]
    lens = (s1 - s0).astype(np.int64)
    assert np.array_equal(lens, q1 - q0)
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    rid = np.repeat(np.arange(len(lens)), lens)
    within = np.arange(int(off[-1])) - off[:-1].astype(np.int64)[rid]
    bases = raw[s0[rid] + within]
    quals = raw[q0[rid] + within]
    hb = raw[: 0]
    hdr_lens = (h1 - h0).astype(np.int64)
    hrid = np.repeat(np.arange(len(lens)), hdr_lens)
    hoff = np.zeros(len(lens) + 1, np.int64)
    hoff[1:] = np.cumsum(hdr_lens)
    hb = raw[h0[hrid] + (np.arange(int(hoff[-1])) - hoff[:-1][hrid])]
    return hb, hoff, bases, quals, off


def round_trip(name, mode, opts, fq, ours_arc, ref_arc, tmp, lossless):
    out = os.path.join(tmp, name + ".ours.fastq")
    p, dt, _ = run([OURS, "decompress", ours_arc, out])
    if p.returncode != 0:
        return {"ok": False, "why": "colord-b200 decompress failed: " + p.stderr[-300:], "decompress_s": dt}
    res = {"decompress_s": dt}
    if lossless:
        same = subprocess.run(["cmp", "-s", fq, out]).returncode == 0
        res.update(ok=same, how="cmp with the input")
    else:
        import numpy as np
        import oracle_lib
        hb, hoff, b, q, off = parse_fastq_arrays(fq)
        hb2, hoff2, b2, q2, off2 = parse_fastq_arrays(out)
        ok = np.array_equal(off, off2) and np.array_equal(b, b2) and np.array_equal(hoff, hoff2) and np.array_equal(hb, hb2)
        how = "bases + headers = input"
        if "none" in opts:      # -q none: every quality is the one representative value (quality_coder.cpp:611-617)
            ok = ok and len(np.unique(q2)) <= 1
            how += "; qualities constant (-q none)"
        else:
            qp = oracle_lib.qual_params(4, [7, 14, 26], 2 if "balanced" in opts else 1)
            want = oracle_lib.qual_lossy(qp, b, q, off)
            ok = ok and np.array_equal(want, q2)
            how += "; qualities = the reference's 4-avg transform (oracle_lib.qual_lossy, pinned on test/*.quan)"
        res.update(ok=bool(ok), how=how)
        del hb, b, q, hb2, b2, q2
    if name == "C1" and ref_arc:      # the reference's own decompressor on its own archive: the two outputs must be the same file
        ref_out = os.path.join(tmp, name + ".ref.fastq")
        p, _, _ = run([REF, "decompress", ref_arc, ref_out])
        same = p.returncode == 0 and subprocess.run(["cmp", "-s", ref_out, out]).returncode == 0
        res["equals_reference_decompress"] = same
        res["ok"] = res["ok"] and same
    try:
        os.remove(out)
    except OSError:
        pass
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4,NS")
    ap.add_argument("--scale", default="C2:0.6,C3:0.125,C4:0.05,NS:0.06", help="fraction of the BASELINE read counts per synthetic configuration (one number = all)")
    ap.add_argument("--runs", type=int, default=1, help="timed runs of each binary (median reported)")
    ap.add_argument("--t1", action="store_true", help="also one `-t 1` run of the reference with its CPU seconds")
    ap.add_argument("--tmp", default="/tmp/ratio_check")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ratio_check.json"))
    ap.add_argument("--md", default=None)
    ap.add_argument("--no-roundtrip", action="store_true")
    ap.add_argument("--ours-opts", default="", help="extra options for colord-b200 (e.g. --compat)")
    args = ap.parse_args()
    os.makedirs(args.tmp, exist_ok=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    nproc = os.cpu_count() or 1
    from colord_b200 import synth
    rows = []
    for name in args.configs.split(","):
        mode, opts, profile, n1, mean_len, cov, lossless = CONFIGS[name]
        fq = os.path.join(args.tmp, name + ".fastq")
        t0 = time.perf_counter()
        if profile is None:
            nbytes, nbases = rebuild_m_bovis(fq)
            n_reads = n1
        else:
            sc = float(dict(x.split(":") for x in args.scale.split(","))[name]) if ":" in args.scale else float(args.scale)
            n_reads = max(200, int(n1 * sc))
            genome_len = int(n_reads * mean_len / cov)
            nbytes, nbases = synth.generate_file(fq, profile, n_reads, genome_len, mean_len, seed=list(CONFIGS).index(name) + 1)
        row = {"config": name, "command": " ".join([mode] + opts), "reads": n_reads, "bases": nbases, "fastq_bytes": nbytes, "generate_s": round(time.perf_counter() - t0, 2), "host_cores": nproc}
        # the stock reference
        ref_arc = os.path.join(args.tmp, name + ".ref.colord")
        ref_t, ref_sz = [], {}
        for _ in range(args.runs):
            p, dt, cpu = run([REF, mode, *opts, "-t", str(nproc), fq, ref_arc], cwd=args.tmp)
            if p.returncode != 0:
                row["reference_error"] = p.stderr[-400:]
                break
            ref_t.append(dt); ref_sz = sizes_of(p.stderr); row["reference_cpu_s"] = round(cpu, 2)
        if ref_t:
            ref_t.sort()
            row.update(reference_archive_bytes=os.path.getsize(ref_arc), reference_streams=ref_sz, reference_wall_s=round(ref_t[len(ref_t) // 2], 3),
                       reference_MBps=round(nbytes / ref_t[len(ref_t) // 2] / 1e6, 2))
        if args.t1 and ref_t:
            p, dt, cpu = run([REF, mode, *opts, "-t", "1", fq, ref_arc + ".t1"], cwd=args.tmp)
            row["reference_t1"] = {"wall_s": round(dt, 3), "cpu_s": round(cpu, 2), "MBps": round(nbytes / dt / 1e6, 2)}
        # ours
        ours_arc = os.path.join(args.tmp, name + ".b200.colord")
        our_t, our_sz = [], {}
        for _ in range(args.runs):
            p, dt, cpu = run([OURS, mode, *opts, "-v", *args.ours_opts.split(), fq, ours_arc], cwd=args.tmp)
            if p.returncode != 0:
                row["ours_error"] = p.stderr[-400:]
                break
            our_t.append(dt); our_sz = sizes_of(p.stderr); row["ours_cpu_s"] = round(cpu, 2)
            row["ours_phases_s"] = {k.strip(): float(v) for k, v in re.findall(r"^  phase (.*): ([0-9.e+-]+) s$", p.stderr, re.M)}
            row["ours_streams_format"] = "compat" if "streams: compat" in p.stderr else "native"
        if our_t:
            our_t.sort()
            row.update(ours_archive_bytes=os.path.getsize(ours_arc), ours_streams=our_sz, ours_wall_s=round(our_t[len(our_t) // 2], 3),
                       ours_file_to_archive_MBps=round(nbytes / our_t[len(our_t) // 2] / 1e6, 2))
        if ref_t and our_t:
            row["archive_ratio"] = round(row["ours_archive_bytes"] / row["reference_archive_bytes"], 5)
            row["stream_ratio"] = {k: round(our_sz[k] / ref_sz[k], 5) for k in ("dna", "quality", "header") if ref_sz.get(k) and k in our_sz}
            row["size_ok"] = row["archive_ratio"] <= 1.005
            row["speedup_file_to_archive"] = round(row["reference_wall_s"] / row["ours_wall_s"], 2)
        if our_t and not args.no_roundtrip:
            try:
                row["round_trip"] = round_trip(name, mode, opts, fq, ours_arc, ref_arc if ref_t else None, args.tmp, lossless)
            except Exception as ex:      # a broken check is a failed check
                row["round_trip"] = {"ok": False, "why": repr(ex)}
        for f in (fq, ref_arc, ref_arc + ".t1", ours_arc):
            try:
                os.remove(f)
            except OSError:
                pass
        rows.append(row)
        print(json.dumps(row), flush=True)
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)
    if args.md:
        with open(args.md, "w") as f:
            f.write("| config | command | bases | FASTQ bytes | reference archive | ours | ratio | dna / quality / header ratio | round trip | reference MB/s (wall) | ours MB/s (file -> archive) |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                sr = r.get("stream_ratio", {})
                f.write(f"| {r['config']} | `{r['command']}` | {r['bases']} | {r['fastq_bytes']} | {r.get('reference_archive_bytes')} | {r.get('ours_archive_bytes')} | "
                        f"{r.get('archive_ratio')} | {sr.get('dna')} / {sr.get('quality')} / {sr.get('header')} | {r.get('round_trip', {}).get('ok')} | "
                        f"{r.get('reference_MBps')} ({r['host_cores']} cores) | {r.get('ours_file_to_archive_MBps')} |\n")


if __name__ == "__main__":
    main()
