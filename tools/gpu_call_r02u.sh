#!/bin/bash
# round 2, call U: sub-warp groups walk back in lockstep, k_anchor_match's writing pass visits only the probes that hit: stage-2 parity
# (+ shards, command line), short bench
set -u
O=gpurun_out/r02u; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_shard.py tests/test_gpu_cli.py -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02u/bench_short.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
