#!/bin/bash
# round 2, call W (2 GPUs): one-pass partition sizes / export of the count table (parity on one GPU first), the count exchange at N = 2,
# and at N = 1 the quality / header coders beside the DNA coder
set -u
O=gpurun_out/r02w; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_shard.py -x -q ) > $O/pytest_stage1.log 2>&1; tail -3 $O/pytest_stage1.log
( time timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -k "multi_gpu" ) > $O/pytest_mgpu.log 2>&1; tail -3 $O/pytest_mgpu.log
BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
grep -E "\[phase\]|\[exchange\]" $O/bench_n2.err | tail -14
BENCH_PHASES=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --stage3-overlap > $O/bench_n1_overlap.json 2> $O/bench_n1_overlap.err
grep -E "\[phase\]" $O/bench_n1_overlap.err | tail -7
python - <<'PY'
import json
for n in ("n2", "n1_overlap"):
    try:
        l = json.loads([x for x in open(f"gpurun_out/r02w/bench_{n}.json") if x.startswith("{")][-1])
        print(n, round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print("ERR", n, e)
PY
tail -3 $O/bench_n1_overlap.err
