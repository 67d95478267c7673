#!/bin/bash
# round 2, call J (8 GPUs): the multi-GPU command line on all GPUs of the box, the bench at N = 8 with its phase trace
set -u
O=gpurun_out/r02j; mkdir -p $O
nvidia-smi -L > $O/smi.txt; nproc >> $O/smi.txt
( time timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "multi_gpu" ) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
python - <<'PY'
import sys; sys.path.insert(0, ".")
from colord_b200 import synth
print(synth.generate_file("/tmp/ns.fastq", "ont", 375000, int(375000 * 8000 / 20.8), 8000, seed=5))
PY
for g in 1 8; do
  if [ $g = 1 ]; then X="--native"; else X="--gpus 8"; fi
  ( time timeout 300 colord_b200/colord-b200 compress-ont $X -v /tmp/ns.fastq /tmp/ns_$g.colord ) > $O/cli_gpus$g.log 2>&1
  grep -E "phase|size|real|rror" $O/cli_gpus$g.log
  ls -l /tmp/ns_$g.colord
done
BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $O/bench_n8.json 2> $O/bench_n8.err
grep "phase\]" $O/bench_n8.err | tail -14; cut -c1-300 $O/bench_n8.json
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02j/bench_n8.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
