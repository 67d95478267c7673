#!/bin/bash
# round 2, call P: half-octave bins + scratch budget from the free memory; thread-per-task backward kernel against the lane-group one
set -u
O=gpurun_out/r02p; mkdir -p $O
for mode in thread group; do
  if [ $mode = group ]; then export CLB_ALIGN_GROUP_BACK=1; fi
  timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_$mode.json 2> $O/bench_$mode.err
  python - $mode <<'PY'
import json, sys
try:
    l = json.loads([x for x in open(f"gpurun_out/r02p/bench_{sys.argv[1]}.json") if x.startswith("{")][-1])
    print(sys.argv[1], round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
done
