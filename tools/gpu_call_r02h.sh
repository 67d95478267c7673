#!/bin/bash
# round 2, call H: the TMA-staged anchor kernel (tests, then the bench), BASELINE configurations 3 and 4 at full size through the device path
set -u
O=gpurun_out/r02h; mkdir -p $O
( CLB_TMA=1 timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_cli.py -x -q -k "not multi_gpu" ) > $O/pytest_tma.log 2>&1; tail -4 $O/pytest_tma.log
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
( CLB_TMA=1 timeout 600 python bench.py $B ) > $O/bench_tma.json 2> $O/bench_tma.err
( timeout 600 python bench.py --config C3 $B ) > $O/bench_c3.json 2> $O/bench_c3.err; tail -2 $O/bench_c3.err
( timeout 600 python bench.py --config C4 --stages 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ) > $O/bench_c4_stage1.json 2> $O/bench_c4_stage1.err; tail -2 $O/bench_c4_stage1.err
( timeout 900 python bench.py --config C4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ) > $O/bench_c4.json 2> $O/bench_c4.err; tail -3 $O/bench_c4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02h/bench_*.json")):
    try:
        l = json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", l["stats"], {k: round(v) for k, v in (l["roofline"] or {}).get("kernel_ms_per_step", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
