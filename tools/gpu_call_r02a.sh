#!/bin/bash
# round 2, call A: run the three code paths round 1 never executed on a device, then bench each
set -u
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/smi.txt; nproc >> $O/smi.txt
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_default.log 2>&1
( CLB_QENC2=1 timeout 600 python -m pytest tests/test_gpu_stage3.py tests/test_gpu_cli.py -m gpu -q ) > $O/pytest_qenc2.log 2>&1
( CLB_EMIT_WARP=1 timeout 600 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_cli.py tests/test_gpu_shard.py -m gpu -q ) > $O/pytest_emitwarp.log 2>&1
( CLB_SLAB_GB=100 timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_slab.log 2>&1
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
( timeout 600 $B ) > $O/bench_default.json 2> $O/bench_default.err
( CLB_QENC2=1 timeout 600 $B ) > $O/bench_qenc2.json 2> $O/bench_qenc2.err
( CLB_EMIT_WARP=1 timeout 600 $B ) > $O/bench_emitwarp.json 2> $O/bench_emitwarp.err
( CLB_SLAB_GB=150 BENCH_PHASES=1 timeout 600 $B ) > $O/bench_slab.json 2> $O/bench_slab.err
tail -3 $O/pytest_*.log
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    l=json.loads([x for x in open(sys.argv[1]) if x.startswith("{")][-1])
    print(l["value"], l["ms_per_step"], {k:round(v) for k,v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e: print("ERR", e)
PY
done
