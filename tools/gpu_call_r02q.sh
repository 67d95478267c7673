#!/bin/bash
# round 2, call Q: DNA walk as a resumable walker (one loop per coder lane), quality encoder fetching the next step's table entries
# ahead + larger chunks: parity (stage 3, compat streams, command line, stage 2), short bench
set -u
O=gpurun_out/r02q; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_exact.py tests/test_gpu_cli.py tests/test_gpu_stage2.py -x -q ) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02q/bench_short.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
CLB_S2_TRACE=1 BENCH_PHASES=1 timeout 600 python bench.py --gbases 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/trace.json 2> $O/trace.err
grep -E "s3q|s3d|\[phase\]" $O/trace.err | tail -22
