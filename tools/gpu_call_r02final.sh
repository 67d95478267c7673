#!/bin/bash
# round 2, final call: the whole GPU suite, the ncu launch list of one bench step at 1 Gbase (time + DRAM traffic per launch) of the
# final build, and the driver's own command (python bench.py)
set -u
O=gpurun_out/r02final; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_1g.csv \
  python bench.py --gbases 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
python - <<'PY'
import csv, collections, json
rows = [r for r in csv.reader(open("gpurun_out/r02final/launches_1g.csv")) if len(r) > 8]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; k, mn, mv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
t, rd, wr, n = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    try:
        name = r[k].split("(")[0].replace("void ", "").replace("clb::", "")[:40]; v = float(r[mv].replace(",", ""))
    except Exception:
        continue
    if r[mn] == "gpu__time_duration.sum": t[name] += v; n[name] += 1
    elif r[mn] == "dram__bytes_read.sum": rd[name] += v
    elif r[mn] == "dram__bytes_write.sum": wr[name] += v
tot = sum(t.values())
print(f"{'kernel':42s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for name, x in t.most_common(40):
    print(f"{name:42s} {n[name]:8d} {x/1e6:10.2f} {100*x/tot:6.1f}% {rd[name]/1e6:11.1f} {wr[name]/1e6:11.1f}")
groups = {"K5-K9 anchors + edit script": ("k_anchor", "k_kmer_anchors", "k_pairs", "k_lis", "k_select", "k_anchor_copy", "k_task", "k_align", "k_decide", "k_read_hist", "k_pending", "k_estimate", "k_emit", "k_node", "k_cview", "k_seg"),
          "K10 DNA entropy": ("k_d_",), "K11 quality entropy": ("k_q_",), "K12 headers": ("k_h_",), "K1+K2 count + threshold": ("k_count", "k_tab", "k_build_surv"), "K3+K4 accepted k-mers + graph": ("k_accept", "k_post", "k_vote", "k_common"), "ingest (k_pack)": ("k_pack", "k_mark")}
out = {}
for g, pre in groups.items():
    names = [x for x in t if any(x.startswith(p) for p in pre)]
    out[g] = {"dram_bytes_per_gbase": sum(rd[x] + wr[x] for x in names), "ms_under_ncu": sum(t[x] for x in names) / 1e6, "launches": sum(n[x] for x in names), "kernels": sorted(names)}
    print(g, round(out[g]["dram_bytes_per_gbase"] / 1e9, 2), "GB per Gbase", round(out[g]["ms_under_ncu"], 1), "ms", out[g]["launches"], "launches")
json.dump(out, open("gpurun_out/r02final/group_traffic.json", "w"), indent=1)
PY
( time timeout 1500 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02final/bench_default.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
    print("e2e", l["e2e"]); print("cpu", l.get("cpu_baseline")); print("ratio", {k: (v if not isinstance(v, dict) else v.get("ratio")) for k, v in l.get("ratio_check", {}).items()})
    print("roofline", {k: v for k, v in l["roofline"].items() if k not in ("kernel_ms_per_step", "groups")})
except Exception as e:
    print("ERR", e)
PY
tail -3 $O/bench_default.err
