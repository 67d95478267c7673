#!/bin/bash
# round 2, call S: per-device slab by default, nodes / candidate views sized once, DNA walk with register windows over the tuple bytes
# and the reference words: whole GPU suite, bench with phase walls (3 timed steps) with and without the slab
set -u
O=gpurun_out/r02s; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for mode in slab noslab; do
  if [ $mode = noslab ]; then export CLB_SLAB_GB=0; fi
  BENCH_PHASES=1 timeout 900 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_$mode.json 2> $O/bench_$mode.err
  grep -E "\[phase\]" $O/bench_$mode.err | tail -21 | awk '{printf "%s=%s ", $3, $4} /readback/ {print ""}'
  python - $mode <<'PY'
import json, sys
try:
    l = json.loads([x for x in open(f"gpurun_out/r02s/bench_{sys.argv[1]}.json") if x.startswith("{")][-1])
    print(sys.argv[1], round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
done
