#!/bin/bash
# round 2, call T: the driver's own command (python bench.py: 3 + 3 steps, end-to-end leg with the quality upload beside stages 1 / 2,
# CPU baseline, ratio check), then the quality / shard / command-line tests that touch clb_append_quals
set -u
O=gpurun_out/r02t; mkdir -p $O
( time timeout 1500 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02t/bench_default.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
    print("e2e", l["e2e"]); print("cpu", l.get("cpu_baseline")); print("ratio", l.get("ratio_check")); print("roofline", {k: v for k, v in l["roofline"].items() if k != "kernel_ms_per_step"})
except Exception as e:
    print("ERR", e)
PY
tail -3 $O/bench_default.err
( time timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_stage3.py -x -q ) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
