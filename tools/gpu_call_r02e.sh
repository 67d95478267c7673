#!/bin/bash
# round 2, call E: -G mode on the device, per-kernel times of the compat coders (ncu launch list), command-line timings with the
# device context made beside the reader
set -u
O=gpurun_out/r02e; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_exact.py -x -q ) > $O/pytest.log 2>&1; tail -12 $O/pytest.log
python - <<'PY'
import sys; sys.path.insert(0, ".")
from colord_b200 import synth
print(synth.generate_file("/tmp/c3s.fastq", "ont", 12500, 5_000_000, 8000, seed=3))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/compat_launches.csv colord_b200/colord-b200 compress-ont --compat -v /tmp/c3s.fastq /tmp/c3s.colord > $O/compat_ncu_run.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02e/compat_launches.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
k, v = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    try: tot[r[k][:60]] += float(r[v].replace(",", "")); cnt[r[k][:60]] += 1
    except Exception: pass
for name, t in tot.most_common(25): print(f"{t/1e6:10.2f} ms  {cnt[name]:6d}  {name}")
PY
timeout 900 python tools/ratio_check.py --configs C3,NS --scale C3:0.0125,NS:0.04 --ours-opts=--compat --no-roundtrip --out $O/ratio_compat.json > $O/ratio_compat.log 2>&1
timeout 900 python tools/ratio_check.py --configs NS --scale NS:0.06 --ours-opts=--native --no-roundtrip --runs 3 --out $O/ratio_native.json > $O/ratio_native.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r02e/ratio_compat.json", "gpurun_out/r02e/ratio_native.json"):
    for r in json.load(open(f)):
        print(r["config"], r["bases"], r.get("ours_streams_format"), "ours", r.get("ours_wall_s"), "ref", r.get("reference_wall_s"), "MB/s", r.get("ours_file_to_archive_MBps"), r.get("ours_phases_s"), r.get("ours_error"))
PY
