#!/bin/bash
# round 2, call V (2 GPUs): k_anchor_match with 512 threads and the one-multiply hash (stage-2 parity first), the multi-GPU command line
# without the slab, the bench at N = 2 with the slab beside torch's exchange tensors and a trace of the count exchange
set -u
O=gpurun_out/r02v; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stage2.py -x -q ) > $O/pytest_stage2.log 2>&1; tail -3 $O/pytest_stage2.log
( time timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -k "multi_gpu" ) > $O/pytest_mgpu.log 2>&1; tail -3 $O/pytest_mgpu.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
grep -E "\[phase\]|\[exchange\]" $O/bench_n2.err | tail -24
python - <<'PY'
import json
for n in ("n1", "n2"):
    try:
        l = json.loads([x for x in open(f"gpurun_out/r02v/bench_{n}.json") if x.startswith("{")][-1])
        print(n, round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print("ERR", n, e)
PY
tail -3 $O/bench_n2.err
