#!/bin/bash
# round 2, call R: where the step's time outside kernels goes at 25 Gbases (wall time per call, stage 2 with event times), and the same
# with the side streams
set -u
O=gpurun_out/r02r; mkdir -p $O
BENCH_PHASES=1 CLB_S2_TRACE=2 timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_phases.json 2> $O/bench_phases.err
grep -E "\[phase\]" $O/bench_phases.err | tail -16
grep -E "^\[s2" $O/bench_phases.err | tail -60
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02r/bench_phases.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
