#!/bin/bash
# round 2, call L: the reworked k_align sweep (anti-diagonal history, shared match vectors, packed carries): stage-2 parity, then the
# whole suite, phase split and a short bench
set -u
O=gpurun_out/r02l; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stage2.py -x -q ) > $O/pytest_stage2.log 2>&1; tail -5 $O/pytest_stage2.log
if grep -q "failed\|error" $O/pytest_stage2.log; then grep -E "Error|assert|FAILED" $O/pytest_stage2.log | head -20; exit 1; fi
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_stage2.py ) > $O/pytest_rest.log 2>&1; tail -4 $O/pytest_rest.log
CLB_LIBRARY=$PWD/colord_b200/libcolord_b200_phases.so timeout 600 python bench.py --gbases 6 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/phases.json 2> $O/phases.err
grep "align phases" $O/phases.err | tail -6
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02l/bench_short.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
