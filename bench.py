#!/usr/bin/env python
"""bench.py — input MB/s of the device hot path on the north-star workload (BASELINE.json).

One "step" = one pass of the hot path built so far over the whole synthetic ONT workload — STAGE 1 (2-bit ingest,
canonical k-mer scan + murmur64 % f filter + count table, thresholding into the filtered set, accepted k-mers per
read, similarity graph with top-c candidates) and STAGE 2 (m-mer anchors against the candidates, edit scripts of the
parts between anchors, edit-script / plain / alternative-read decisions, CompactES tuples).  Stage 3 (entropy coders) is
all three streams in native containers: the quality stream (the reference's lossy 4-avg transform + context model, static
tables, interleaved rANS), the DNA-tuple stream and the header stream (the reference's event models, static tables, 64 range-
coder lanes per pack); `config.stages` says so and the reference arm times the SAME stages of the reference
(`--stages 1` / `12` / `12q` / `12qd` restrict both arms).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--gbases G] [--impl reference]

value  = FASTQ-equivalent input bytes of the whole job / step time, inputs resident in HBM.
e2e    = same through the C-ABI with HOST (pinned) buffers: H2D inside the timed region, candidates read back.
Multi-GPU (torchrun): reads shard by id (strong scaling); one all-to-all + one all-gather of k-mer tables, one all-gather of the
reference reads (every rank keeps those of the ranks before it as context reads: same candidates and tuples as one GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# north-star workload (SURVEY.md §8d "NS"): compress-ont default on a 50 GB ONT FASTQ, mean read 8 kb
NS = dict(k=24, modulo=12, min_count=4, max_count=80, max_candidates=5, sparse_g=1.0, sparse_exponent=1.0,
          mean_len=8000, genome_len=1_200_000_000, err=(0.04, 0.03, 0.03))
# stage 2 at the same preset (arg_parse.cpp:154 "memory": level 1) and the anchor length compression.cpp:57-94 picks for 25 Gbases
NS_S2 = dict(anchor_len=22, min_part_len_alt=64, max_recurence=3, min_anchors=1, min_mmer_frac=0.5, min_mmer_force=0.9,
             max_matches_mult=10.0, es_cost_mult=1.0)
# --config: the other BASELINE configurations that run through the same device path (presets of arg_parse.cpp:89-408,
# k / anchor length by compression.cpp:41-93 for the configuration's size).  Each overrides the workload dictionaries above.
CONFIGS = {
    "NS": dict(gbases=25.0, cmd="compress-ont", cli=[], level=1, qual="4-avg", sparse=True, text="compress-ont default"),
    "C3": dict(gbases=8.0, cmd="compress-ont", cli=["-p", "balanced"], level=2, qual="4-avg", sparse=True, text="compress-ont -p balanced (BASELINE config 3: 1 M ONT reads x 8 kb)",
               ns=dict(k=23, modulo=9, min_count=3, max_count=100, max_candidates=8, sparse_g=2.0, genome_len=400_000_000), s2=dict(anchor_len=21, min_part_len_alt=48, max_recurence=5)),
    "C4": dict(gbases=20.0, cmd="compress-pbraw", cli=["-q", "none", "-p", "ratio"], level=3, qual=None, sparse=False, text="compress-pbraw -q none -p ratio (BASELINE config 4: 2 M CLR subreads x 10 kb)",
               ns=dict(k=24, modulo=8, min_count=2, max_count=120, max_candidates=10, mean_len=10000, genome_len=1_000_000_000, err=(0.02, 0.04, 0.07)), s2=dict(anchor_len=22, min_part_len_alt=48, max_recurence=6)),
}
CFG = dict(CONFIGS["NS"], name="NS")


def apply_config(name):
    c = CONFIGS[name]
    NS.update(c.get("ns", {}))
    NS_S2.update(c.get("s2", {}))
    CFG.clear(); CFG.update(c, name=name)


HEADER_BYTES = 46 + 6          # "@read_<i> ch=<n> start_time=<ISO>\n" + "\n+\n" + two line ends


def fastq_bytes(n_bases, n_reads):
    return 2 * n_bases + HEADER_BYTES * n_reads


# ----------------------------------------------------------------------------------------------------
# synthetic reads generated on the device (plumbing; same model as colord_b200.synth / BASELINE.md §2)
# ----------------------------------------------------------------------------------------------------
def make_genome(torch, device, genome_len, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 4, (genome_len,), dtype=torch.uint8, device=device, generator=g)


def gen_reads(torch, device, genome, read_lo, read_hi, seed, out_bases=None):
    """ASCII bases + offsets of reads [read_lo, read_hi) of the workload.  Returns (bases u8, offsets i64)."""
    G = genome.numel()
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    e_sub, e_del, e_ins = NS["err"]
    chunks, lens_all = [], []
    CH = 20000
    for lo in range(read_lo, read_hi, CH):
        hi = min(read_hi, lo + CH)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + lo)
        n = hi - lo
        u = torch.rand((2, n), device=device, generator=g).clamp_min(1e-12)
        lens = (-(NS["mean_len"] / 2.0) * (u[0].log() + u[1].log())).clamp(200, 200000).to(torch.int64)
        start = (torch.rand(n, device=device, generator=g) * (G - 2.5 * lens.double() - 16).clamp_min(1)).to(torch.int64)
        rev = torch.rand(n, device=device, generator=g) < 0.5
        offs = torch.zeros(n + 1, dtype=torch.int64, device=device)
        offs[1:] = torch.cumsum(lens, 0)
        total = int(offs[-1])
        rid = torch.repeat_interleave(torch.arange(n, device=device), lens, output_size=total)
        r = torch.rand(total, device=device, generator=g)
        is_sub = r < e_sub
        is_del = (r >= e_sub) & (r < e_sub + e_del)
        is_ins = (r >= e_sub + e_del) & (r < e_sub + e_del + e_ins)
        step = 1 + is_del.to(torch.int64) - is_ins.to(torch.int64)
        cs = torch.cumsum(step, 0)
        src_off = cs - cs[offs[:-1]][rid] + step[offs[:-1]][rid] - step      # exclusive scan restarted per read
        span = (2 * lens + 8)[rid]
        src = torch.where(rev[rid], start[rid] + span - 1 - src_off, start[rid] + src_off).clamp_(0, G - 1)
        base = genome[src]
        base = torch.where(rev[rid], 3 - base, base)
        rnd = torch.randint(0, 4, (total,), dtype=torch.uint8, device=device, generator=g)
        base = torch.where(is_sub, (base + 1 + rnd % 3) & 3, base)
        base = torch.where(is_ins, rnd, base)
        chunks.append(lut[base.long()])
        lens_all.append(lens)
        del rid, r, is_sub, is_del, is_ins, step, cs, src_off, span, src, base, rnd
    lens = torch.cat(lens_all)
    offsets = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(lens, 0)
    bases = torch.cat(chunks) if len(chunks) > 1 else chunks[0]
    return bases, offsets


def sparse_range(stats, p):
    # compression.cpp:443, :501-503
    mean_read_len = int(stats["tot_kmers"] * p["modulo"] / max(1, stats["n_reads"]) + p["k"] - 1)
    return max(1, int(p["sparse_g"] * stats["n_unique_counted"] * p["modulo"] / max(1, mean_read_len)))


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_SMI_MS", "1000")],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the STOCK reference binary (oracle/_ref/colord, the unmodified CLI built by oracle/Makefile) on a
# FASTQ file that is a bounded sample of the workload — wall time from process start to exit, all host threads
# ----------------------------------------------------------------------------------------------------
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "colord")
OUR_CLI = os.path.join(ROOT, "colord_b200", "colord-b200")
SIZE_KEYS = (("DNA", "dna"), ("Quality", "quality"), ("Header", "header"))


def sample_fastq(path, n_reads):
    """n_reads synthetic ONT reads of the north-star workload (same error / quality / header model, 20.8x coverage) as a FASTQ file"""
    from colord_b200 import synth
    nbytes, nbases = synth.generate_file(path, "clr" if CFG["cmd"] == "compress-pbraw" else "ont", n_reads, max(100_000, int(n_reads * NS["mean_len"] / 20.8)), NS["mean_len"], seed=1)
    return nbytes, nbases


def run_cli(exe, fastq, out, extra=()):
    """One `compress-ont` run at the north star's preset (k / anchor length forced to what a 50 GB input selects).  -> (wall s, stream sizes)"""
    import re
    cmd = [exe, CFG["cmd"], *CFG["cli"], "-k", str(NS["k"]), "-a", str(NS_S2["anchor_len"]), *extra, fastq, out]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=3600, cwd=os.path.dirname(out))
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{os.path.basename(exe)} failed: {r.stderr[-300:]}")
    sizes = {k: int(m.group(1)) for name, k in SIZE_KEYS for m in [re.search(rf"^{name} size\s*:\s*(\d+)", r.stderr, re.M)] if m}
    m = re.search(r"phase device context \(rest of it\): ([0-9.e+-]+) s", r.stderr)      # colord-b200 -v: what the process waited for the CUDA start-up
    if m:
        sizes["cuda_startup_wait_s"] = float(m.group(1))
    return dt, sizes


def run_port_stage1(n_reads=2000):
    """Fallback CPU baseline where the reference binary is missing: the scalar C oracle (single thread), stage 1 only."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib
    from colord_b200 import synth
    s = synth.generate(n_reads, 1_000_000, NS["mean_len"], seed=1, profile="ont")
    t0 = time.time()
    km, ct, st = oracle_lib.count_kmers(s.bases, s.offsets, NS["k"], NS["modulo"], NS["min_count"], NS["max_count"])
    off, acc = oracle_lib.accepted_kmers(s.bases, s.offsets, NS["k"], NS["modulo"], km)
    rng = max(1, int(st["n_unique_counted"] * NS["modulo"] / NS["mean_len"]))
    sampled = oracle_lib.sampler(rng, 1.0, 0, s.n_reads)
    oracle_lib.sim_graph(off, acc, np.zeros(s.n_reads, np.uint8), sampled, NS["max_candidates"], NS["max_count"])
    return s.fastq_bytes(), time.time() - t0


def cpu_baseline(sample_reads=125000, with_ratio_check=True):
    """The stock reference on a 2 GB sample of the workload, once (~20 s on 16 cores); beside it the command line of this repo on the SAME file
    (ratio_check: archive sizes of both, and file -> archive MB/s of colord-b200 in both stream formats)."""
    cores = os.cpu_count() or 1
    if not os.path.exists(REF_BIN):
        nbytes, dt = run_port_stage1()
        return {"value": nbytes / dt / 1e6, "unit": "MB/s", "cores": 1, "kind": "port", "sample": "2 000 synthetic ONT reads; oracle/stage1.c scalar port, stage 1 only"}, None
    with tempfile.TemporaryDirectory() as tmp:
        fq = os.path.join(tmp, "sample.fastq")
        nbytes, nbases = sample_fastq(fq, sample_reads)
        dt, ref_sizes = run_cli(REF_BIN, fq, os.path.join(tmp, "ref.colord"), ["-t", str(cores)])
        ref_bytes = os.path.getsize(os.path.join(tmp, "ref.colord"))
        base = {"value": nbytes / dt / 1e6, "unit": "MB/s", "cores": cores, "kind": "reference", "wall_s": dt,
                "sample": f"{sample_reads} synthetic ONT reads of the workload, {nbases} bases, {nbytes} FASTQ bytes; unmodified `colord {CFG['cmd']} {' '.join(CFG['cli'])} -k {NS['k']} -a {NS_S2['anchor_len']} -t {cores}`, "
                          "wall time process start -> exit, one run"}
        check = None
        if with_ratio_check and os.path.exists(OUR_CLI):
            check = {"config": f"{CFG['text']} on the same {nbytes}-byte FASTQ file", "ref_bytes": ref_bytes, "ref_streams": ref_sizes}
            for fmt in ("native", "compat"):
                try:
                    t, sizes = run_cli(OUR_CLI, fq, os.path.join(tmp, fmt + ".colord"), ["--" + fmt, "-v"])
                    ours = os.path.getsize(os.path.join(tmp, fmt + ".colord"))
                    check[fmt] = {"ours_bytes": ours, "ratio": ours / ref_bytes, "streams": sizes, "file_to_archive_MBps": nbytes / t / 1e6, "wall_s": t}
                except Exception as ex:
                    check[fmt] = {"error": str(ex)[:300]}
            check["ours_bytes"] = check.get("native", {}).get("ours_bytes")
        return base, check


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_reads = int(os.environ.get("BENCH_REF_READS", "62500"))      # 0.5 Gbases = 1 GB of FASTQ per step: ~10 s on 16 cores
    kind = "reference" if os.path.exists(REF_BIN) else "port"
    times, sizes, nbytes, nbases = [], {}, 0, 0
    with tempfile.TemporaryDirectory() as tmp:
        if kind == "reference":
            fq = os.path.join(tmp, "sample.fastq")
            nbytes, nbases = sample_fastq(fq, n_reads)
        for i in range(args.warmup + args.steps):
            if kind == "reference":
                dt, sizes = run_cli(REF_BIN, fq, os.path.join(tmp, "ref.colord"), ["-t", str(cores)])
            else:
                nbytes, dt = run_port_stage1()
            if i >= args.warmup:
                times.append(dt)
        archive = os.path.getsize(os.path.join(tmp, "ref.colord")) if kind == "reference" else None
    ms = 1e3 * sum(times) / len(times)
    v = nbytes / (ms / 1e3) / 1e6
    sample = (f"{n_reads} synthetic ONT reads of the workload / {nbases} bases / {nbytes} FASTQ bytes per step (bounded sample); the unmodified reference CLI "
              f"`colord {CFG['cmd']} {' '.join(CFG['cli'])} -k {NS['k']} -a {NS_S2['anchor_len']} -t {cores}` (stock code path, file in -> archive out, wall time process start -> exit)"
              if kind == "reference" else "oracle/stage1.c scalar port, stage 1 only (the reference binary is not built here)")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": "MB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
        "archive_bytes": archive, "stream_bytes": sizes,
        "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def metric_name(args):
    return f"input MB/s, {CFG['text']}, " + {
        "12qdh": "stages 1+2+3 (k-mer filter, similarity graph, anchors + edit scripts -> tuples, DNA-tuple, 4-avg quality and header entropy coders)",
        "12qd": "stages 1+2+3 without headers (k-mer filter, similarity graph, anchors + edit scripts -> tuples, DNA-tuple and 4-avg quality entropy coders)",
        "12q": "stages 1+2 + quality stream of stage 3 (k-mer filter, similarity graph, anchors + edit scripts -> tuples, 4-avg quality coder)",
        "12": "stages 1+2 (k-mer filter, similarity graph, anchors + edit scripts -> tuples)",
        "1": "stage 1 (k-mer filter + similarity graph)"}[args.stages]


def workload_config(args, n_reads):
    return {"workload": f"{CFG['text']} (k{NS['k']} f{NS['modulo']} L{NS['min_count']} H{NS['max_count']} c{NS['max_candidates']} {'sparse g=%g' % NS['sparse_g'] if CFG['sparse'] else 'all reads are references'}), "
                        f"synthetic FASTQ ~{2 * args.gbases:.0f} GB ({args.gbases:g} Gbases, mean read {NS['mean_len'] / 1000:g} kb, genome {NS['genome_len'] * args.gbases / CFG['gbases'] / 1e9:.3g} Gb, {100 * sum(NS['err']):.0f}% errors)",
            "stages": ("stages 1+2 (1a count+filter, 1b accepted k-mers + similarity graph, 2 anchors/edit scripts/decisions/CompactES tuples; a%d lvl%d)" % (NS_S2["anchor_len"], CFG["level"])
                       + {"12qdh": "; stage 3: DNA-tuple stream (level 1) + quality stream (4-avg, thresholds 7 14 26) + header stream in native containers",
                          "12qd": "; stage 3: DNA-tuple stream (level 1) + quality stream (4-avg, thresholds 7 14 26) in native containers; header coder not on device yet",
                          "12q": "; stage 3: quality stream (4-avg, thresholds 7 14 26) only", "12": "; stage 3 not included"}[args.stages])
                      if args.stages != "1" else "stage 1 only (1a count+filter, 1b accepted k-mers + similarity graph)",
            "n_reads": n_reads, "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"reads sharded by id over {args.gpus} GPU(s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="NS", choices=sorted(CONFIGS), help="workload: the north star (default) or BASELINE configuration 3 / 4 at its own preset and size")
    ap.add_argument("--gbases", type=float, default=None, help="workload size in Gbases (default: the configuration's own — north star 25 = 50 GB FASTQ, C3 8, C4 20)")
    ap.add_argument("--stages", default="12qdh", choices=["1", "12", "12q", "12qd", "12qdh"], help="hot-path stages inside a step (both arms); q / d / h = quality / DNA-tuple / header stream of stage 3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--side-streams", action="store_true", help="code the quality / header streams in a second host thread beside stage 2 instead of after it "
                    "(measured on B200, 25 Gbases: 12.5-13.3 s per step against 11.8-12.0 s serial - the issue-bound quality kernels slow the latency-bound alignment down)")
    args = ap.parse_args()
    apply_config(args.config)
    ref_gbases = CFG["gbases"]
    if args.gbases is None:
        args.gbases = ref_gbases
    if args.impl == "reference":
        return main_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from colord_b200 import lib
    from colord_b200.dist import exchange_counts_and_finalize, exchange_reference_reads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:      # the library's device slab (csrc/slab.h) takes the free memory minus a reserve when a context is created: leave room
        # for the tensors torch allocates for the two exchanges of dist.py (pairs sent + received ~ 48 GB / world, reference reads)
        os.environ.setdefault("CLB_SLAB_RESERVE_GB", str(10 + 56 // world))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; colord_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib.load()

    p = NS
    n_reads_total = int(args.gbases * 1e9 / p["mean_len"])
    lo, hi = rank * n_reads_total // world, (rank + 1) * n_reads_total // world
    # reduced runs (--gbases < 25) keep the north star's 20.8x coverage by shrinking the genome with the workload
    genome_len = max(100_000, int(p["genome_len"] * args.gbases / ref_gbases))
    genome = make_genome(torch, device, genome_len, seed=1234)
    bases, offsets = gen_reads(torch, device, genome, lo, hi, seed=99)
    del genome
    torch.cuda.empty_cache()
    quals = None
    if args.stages in ("12q", "12qd", "12qdh") and CFG["qual"]:      # phred ~ clip(N(12, 5), 1, 40) + 33, generated in slices
        quals = torch.empty(bases.numel(), dtype=torch.uint8, device=device)
        gq = torch.Generator(device=device)
        gq.manual_seed(4242 + rank)
        for q0 in range(0, bases.numel(), 1 << 28):
            q1 = min(bases.numel(), q0 + (1 << 28))
            quals[q0:q1] = (torch.randn(q1 - q0, device=device, generator=gq) * 5 + 12).round_().clamp_(1, 40).to(torch.uint8) + 33
        torch.cuda.empty_cache()
    hdr_host = hdr_off_host = hdr_dev = hdr_off_dev = None
    if args.stages == "12qdh":      # "@read_<i> ch=<1..512> start_time=<ISO>" (same text as colord_b200.synth), global read index
        rng_h = np.random.default_rng(777 + rank)
        chs = rng_h.integers(1, 513, hi - lo)
        hl = [b"@read_%d ch=%d start_time=2020-01-01T%02d:%02d:%02dZ" % (i, c, (i // 3600) % 24, (i // 60) % 60, i % 60) for i, c in zip(range(lo, hi), chs.tolist())]
        hdr_off_host = np.zeros(hi - lo + 1, np.uint64)
        hdr_off_host[1:] = np.cumsum(np.fromiter((len(h) for h in hl), np.int64, len(hl)))
        hdr_host = torch.empty(int(hdr_off_host[-1]), dtype=torch.uint8, pin_memory=True)
        hdr_host.numpy()[:] = np.frombuffer(b"".join(hl), np.uint8)
        del hl, chs
        hdr_dev = hdr_host.to(device)
        hdr_off_dev = torch.from_numpy(hdr_off_host.view(np.int64)).to(device)
    free_b, total_b = torch.cuda.mem_get_info()
    print(f"[bench] rank {rank}: inputs resident, {free_b / 2**30:.1f} of {total_b / 2**30:.1f} GiB free "
          f"(torch reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB)", file=sys.stderr)
    n_local = hi - lo
    n_bases_local = int(bases.numel())
    off_u64 = offsets.contiguous()
    tot = torch.tensor([n_bases_local, n_local], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(tot)
    n_bases_total, n_reads_all = int(tot[0]), int(tot[1])
    job_bytes = fastq_bytes(n_bases_total, n_reads_all)      # header lines counted at their nominal 46 bytes (same formula for every --stages)

    stream = torch.cuda.Stream(device=device)
    peak, peak_src = measured_peak_hbm()
    lens_host = np.diff(offsets.cpu().numpy())
    out_pinned = {}      # e2e: the finished streams land in pinned host buffers (sized on first use, kept across steps)

    def read_stream(ctx, which):
        n = {"dna": lambda: ctx.dna_stream_size(), "qual": ctx.qual_size, "hdr": lambda: ctx.hdr_stream_size()}[which]()
        if which not in out_pinned or out_pinned[which].numel() < n:
            out_pinned[which] = torch.empty(int(n * 1.1) + 4096, dtype=torch.uint8, pin_memory=True)
        ctx.stream_into(which, out_pinned[which].data_ptr(), out_pinned[which].numel())
        return out_pinned[which].numpy()[:n]

    phases_on = os.environ.get("BENCH_PHASES") is not None

    def one_step(host_bases=None, host_offsets=None, host_quals=None, host_hdr=None, profile=False, readback=False):
        t_ph = [time.perf_counter()]

        def phase(name):      # BENCH_PHASES=1: wall time of every call of the step (debugging aid; the calls synchronise by themselves)
            if phases_on and rank == 0:
                t = time.perf_counter()
                print(f"[phase] {'e2e' if host_bases is not None else 'res'} {name:14s} {1e3 * (t - t_ph[0]):9.1f} ms   free {torch.cuda.mem_get_info()[0] / 2**30:6.1f} GiB", file=sys.stderr)
                t_ph[0] = t
        ctx = lib.Context(p["k"], p["modulo"], p["min_count"], p["max_count"], p["max_candidates"], expected_bases=n_bases_local, device=local_rank)
        ctx.set_stream(stream.cuda_stream)
        if profile:
            ctx.profile_enable(True)
        qual_up, qual_up_err = None, []
        if host_bases is None:
            ctx.append_reads_device(bases.data_ptr(), off_u64.data_ptr(), n_local)
        else:
            ctx.append_reads(host_bases, host_offsets)
            if host_quals is not None and CFG["qual"]:
                # end to end: the qualities travel to the device (clb_append_quals, stage-3 stream) from a second host thread while
                # stages 1 and 2 run — the way the command line streams them beside the bases; 1 GiB per call
                def qual_upload():
                    try:
                        for a in range(0, len(host_quals), 1 << 30):
                            ctx.append_quals(host_quals[a:a + (1 << 30)])
                    except Exception as ex:
                        qual_up_err.append(ex)
                qual_up = threading.Thread(target=qual_upload)
                qual_up.start()
        phase("append")
        with torch.cuda.stream(stream):
            stats = exchange_counts_and_finalize(ctx, device, n_local)
        phase("finalize")
        rng = sparse_range(stats, p)
        sampled = lib.sampler(rng, p["sparse_exponent"], 0, n_reads_all)[lo:hi] if CFG["sparse"] else np.ones(hi - lo, np.uint8)
        if world > 1:      # global reference-read set: the reference reads of the shards before mine become my context reads
            with torch.cuda.stream(stream):
                exchange_reference_reads(ctx, device, sampled, lens_host)
                stream.synchronize()
        # the quality and header streams do not depend on stage 2 (level 1): with --side-streams a second host thread codes them on
        # the context's stage-3 stream while the main thread runs the graph, stage 2 and the DNA stream (the reference gives these
        # coders their own threads as well, compression.cpp:654-689); default: after stage 2, which measured faster
        def side_streams():
            if args.stages == "12qdh":
                if host_hdr is None:
                    ctx.hdr_encode(bytes_=hdr_dev.data_ptr(), offsets=hdr_off_dev.data_ptr(), n=n_local, on_device=True)
                else:
                    ctx.hdr_encode(bytes_=host_hdr[0], offsets=host_hdr[1])
            if not CFG["qual"]:
                return
            if host_quals is None:
                ctx.qual_encode(4, [7, 14, 26], CFG["level"], quals.data_ptr(), off_u64.data_ptr(), on_device=True)
            else:
                qual_up.join()
                if qual_up_err:
                    raise qual_up_err[0]
                ctx.qual_encode(4, [7, 14, 26], CFG["level"], None, None)
        side, side_err = None, []
        if args.stages in ("12q", "12qd", "12qdh") and args.side_streams:
            def side_main():
                try:
                    side_streams()
                except Exception as ex:      # re-raised on the main thread
                    side_err.append(ex)
            side = threading.Thread(target=side_main)
            side.start()
        ctx.graph_build(sampled)
        phase("graph")
        out = None
        if args.stages in ("12", "12q", "12qd", "12qdh"):
            ctx.encode(lib.EncodeParams(*[NS_S2[k] for k in ("anchor_len", "min_part_len_alt", "max_recurence", "min_anchors",
                                                           "min_mmer_frac", "min_mmer_force", "max_matches_mult", "es_cost_mult")]))
            phase("encode")
            if args.stages in ("12qd", "12qdh"):
                ctx.dna_encode(CFG["level"])
            phase("dna")
            if side is not None:
                side.join()
                if side_err:
                    raise side_err[0]
            elif args.stages in ("12q", "12qd", "12qdh"):
                side_streams()
            phase("hdr+qual")
            if readback:           # what leaves the device: the finished streams (the tuples too while the DNA coder is not included)
                if args.stages == "12qdh":
                    out = (read_stream(ctx, "dna"),) + ((read_stream(ctx, "qual"),) if CFG["qual"] else ()) + (read_stream(ctx, "hdr"),)
                elif args.stages == "12qd":
                    out = (read_stream(ctx, "dna"), read_stream(ctx, "qual"))
                else:
                    out = ctx.encoded(ctx.n_reads)
                    if args.stages == "12q":
                        out = out + (ctx.qual_stream(),)
        elif readback:
            out = ctx.graph_candidates()
        ctx.synchronize()
        phase("readback")
        return ctx, stats, out

    def timed(n_warm, n_steps, **kw):
        times, last = [], None
        prof = {}
        launches = 0
        for i in range(n_warm + n_steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx, stats, out = one_step(profile=(i >= n_warm), **kw)
            e1.record(stream)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if i >= n_warm:
                times.append(e0.elapsed_time(e1))
                for k, (ms, n) in ctx.profile().items():
                    a = prof.get(k, (0.0, 0))
                    prof[k] = (a[0] + ms, a[1] + n)
                launches += ctx.kernel_launches
            last = (stats, out)
            ctx.close()
        t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), prof, launches // max(1, n_steps), last

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, prof, launches, (stats, _) = timed(args.warmup, args.steps)
    value = job_bytes / (ms / 1e3) / 1e6

    e2e = None
    if not args.no_e2e:
        host_bases = torch.empty(n_bases_local, dtype=torch.uint8, pin_memory=True)
        host_bases.copy_(bases)
        host_off = offsets.cpu().numpy().astype(np.uint64)
        hb = host_bases.numpy()
        hq = None
        if quals is not None:
            host_quals = torch.empty(n_bases_local, dtype=torch.uint8, pin_memory=True)
            host_quals.copy_(quals)
            hq = host_quals.numpy()
        hh = None if hdr_host is None else (hdr_host.numpy(), hdr_off_host)
        # the device-resident copies of the inputs are not part of the end-to-end path: give their memory back first
        del bases, quals, hdr_dev
        bases = quals = hdr_dev = None
        torch.cuda.empty_cache()
        e_ms, _, _, (_, out) = timed(1, max(1, min(args.steps, 2)), host_bases=hb, host_offsets=host_off, host_quals=hq, host_hdr=hh, readback=True)
        d2h = int(sum(x.nbytes for x in out))
        sizes_all = torch.tensor([int(x.nbytes) for x in out], dtype=torch.int64, device=device)
        if world > 1:      # whole-job stream sizes (the ratio check across shard counts)
            dist.all_reduce(sizes_all)
        e2e = {"value": job_bytes / (e_ms / 1e3) / 1e6, "unit": "MB/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(n_bases_local * (2 if hq is not None else 1) + host_off.nbytes + (0 if hh is None else hh[0].nbytes + hh[1].nbytes)),
               "d2h_bytes_per_step": d2h, "host_memory": "pinned", "stream_bytes_rank0": [int(x.nbytes) for x in out], "stream_bytes_job": sizes_all.tolist()}
        del host_bases, hb
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        # roofline of the dominant kernel GROUP of the step: ALGORITHMIC bytes per base of SURVEY.md §8(d) (restated in DESIGN.md §4)
        # x the bases of the job / the group's device time (CUDA events on the launching streams, summed over its launches)
        f, cnd = float(p["modulo"]), float(p["max_candidates"])
        groups = {      # name: (kernel classes, algorithmic bytes per input base)
            "ingest (k_pack)": (["k_pack"], 0.25 + 1.0),
            "K1+K2 count + threshold": (["k_count", "k_tab_misc", "k_finalize"], 0.25 + 16.0 / f),
            "K3+K4 accepted k-mers + graph": (["k_accept", "k_postings", "k_vote", "k_common"], 0.25 + 4.0 / f + 12.0 / f),
            "K5-K9 anchors + edit script": (["k_anchors", "k_align", "k_encode", "k_decide", "k_estimate", "k_emit"], 0.25 * (1 + 2 * cnd) + 1.25),
            "K10 DNA entropy": (["k_dna"], 1.25 + 0.25 + 0.25),
            "K11 quality entropy": (["k_qual"], 1.0 + 0.25 + 0.25),
            "K12 headers": (["k_hdr"], 0.03),
        }
        per_step = {k: v[0] / max(1, args.steps) for k, v in prof.items() if v[1]}
        roof = None
        if per_step:
            g_ms = {g: sum(per_step.get(k, 0.0) for k in ks) for g, (ks, _) in groups.items()}
            g_n = {g: sum(prof.get(k, (0, 0))[1] for k in ks) // max(1, args.steps) for g, (ks, _) in groups.items()}
            top = max(g_ms, key=lambda g: g_ms[g])
            alg_b = groups[top][1]
            achieved = alg_b * n_bases_local / (g_ms[top] / 1e3) / 1e9
            traffic = None
            try:      # DRAM bytes of the group from an ncu pass (profiles/r02_group_traffic.json: bytes per Gbase of input), per launch like `achieved`
                with open(os.path.join(ROOT, "profiles", "r02_group_traffic.json")) as fh:
                    traffic = json.load(fh)[top]["dram_bytes_per_gbase"] * (n_bases_local / 1e9) / max(1, g_n[top])
            except Exception:
                pass
            roof = {"kernel": top, "kernels": groups[top][0], "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "launches": g_n[top], "ms_per_launch": g_ms[top] / max(1, g_n[top]),
                    "algorithmic_bytes_per_base": alg_b, "algorithmic_bytes_per_launch": alg_b * n_bases_local / max(1, g_n[top]),
                    "groups": {g: {"ms_per_step": g_ms[g], "algorithmic_bytes_per_base": groups[g][1], "achieved_GBps": groups[g][1] * n_bases_local / (g_ms[g] / 1e3) / 1e9 if g_ms[g] else None,
                                   "frac": groups[g][1] * n_bases_local / (g_ms[g] / 1e3) / 1e9 / peak if g_ms[g] else None} for g in groups},
                    "whole_path": {"algorithmic_bytes_per_base": sum(v[1] for v in groups.values()), "achieved_GBps": sum(v[1] for v in groups.values()) * n_bases_local / (ms / 1e3) / 1e9,
                                   "frac": sum(v[1] for v in groups.values()) * n_bases_local / (ms / 1e3) / 1e9 / peak},
                    "note": "the edit-script group is bound by the dependent-issue latency of the bit-vector recurrence and by divergent walks, not by HBM (profiles/)",
                    "kernel_ms_per_step": per_step}
        line = {
            "metric": metric_name(args), "value": value, "unit": "MB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic (generated on device, BASELINE.md §2 error model)",
            "config": workload_config(args, n_reads_all), "job_fastq_bytes": job_bytes, "stats": stats,
            "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], line["ratio_check"] = cpu_baseline()
            except Exception as ex:      # keep the GPU numbers even if the host baseline cannot run
                line["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
