#!/bin/bash
# round 2, call G: the whole GPU suite, compat timings after the warp-per-pack range coder, the default bench run
set -u
O=gpurun_out/r02g; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -8 $O/pytest.log
timeout 900 python tools/ratio_check.py --configs C1,C3,NS --scale C3:0.0125,NS:0.04 --ours-opts=--compat --no-roundtrip --out $O/ratio_compat.json > $O/ratio_compat.log 2>&1
python - <<'PY'
import json
for r in json.load(open("gpurun_out/r02g/ratio_compat.json")):
    print(r["config"], r["bases"], r.get("ours_streams_format"), "ours", r.get("ours_wall_s"), "ref", r.get("reference_wall_s"), "ratio", r.get("archive_ratio"), r.get("ours_phases_s"), r.get("ours_error"))
PY
( time timeout 1500 python bench.py ) > $O/bench.json 2> $O/bench.err; tail -4 $O/bench.err; cut -c1-600 $O/bench.json
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
