#!/bin/bash
# round 2, call D: whole GPU suite with the compat + streaming paths, compat / native timings of the command line, one short bench
set -u
O=gpurun_out/r02d; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -12 $O/pytest.log
timeout 900 python tools/ratio_check.py --configs C1,C3,NS --scale C3:0.0125,NS:0.04 --ours-opts=--compat --no-roundtrip --out $O/ratio_compat.json --md $O/ratio_compat.md > $O/ratio_compat.log 2>&1; cat $O/ratio_compat.md
timeout 900 python tools/ratio_check.py --configs C3,NS --scale C3:0.125,NS:0.06 --ours-opts=--native --no-roundtrip --out $O/ratio_native.json --md $O/ratio_native.md > $O/ratio_native.log 2>&1; cat $O/ratio_native.md
python - <<'PY'
import json
for f in ("gpurun_out/r02d/ratio_compat.json", "gpurun_out/r02d/ratio_native.json"):
    for r in json.load(open(f)):
        print(r["config"], r["bases"], r.get("ours_streams_format"), "ours", r.get("ours_wall_s"), "ref", r.get("reference_wall_s"), r.get("ours_phases_s"), r.get("ours_error"))
PY
( timeout 900 python bench.py --steps 2 --warmup 1 ) > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; cut -c1-1500 $O/bench.json
