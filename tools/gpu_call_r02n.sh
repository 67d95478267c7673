#!/bin/bash
# round 2, call N: backward half of k_align as one thread per task (cp.async ring, streaming canonical form); DNA walk with the
# event queue: parity (stage 2, stage 3, shards, command line), short bench, synchronising stage trace at 6 Gbases
set -u
O=gpurun_out/r02n; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_shard.py -x -q ) > $O/pytest_stage2.log 2>&1; tail -5 $O/pytest_stage2.log
if grep -q "failed\|error" $O/pytest_stage2.log; then grep -E "Error|assert|FAILED" $O/pytest_stage2.log | head -20; fi
( time timeout 900 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_cli.py -x -q ) > $O/pytest_stage3.log 2>&1; tail -5 $O/pytest_stage3.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02n/bench_short.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
CLB_S2_TRACE=1 BENCH_PHASES=1 timeout 600 python bench.py --gbases 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/trace.json 2> $O/trace.err
grep -E "s3q|s3d|\[phase\]" $O/trace.err | tail -22
grep -E "^\[s2\]|\[s2 " $O/trace.err | tail -40
