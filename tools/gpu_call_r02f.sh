#!/bin/bash
# round 2, call F (2 GPUs): the multi-GPU command line (NCCL exchanges in C++), timings against one GPU, a short 2-GPU bench with phases
set -u
O=gpurun_out/r02f; mkdir -p $O
nvidia-smi -L > $O/smi.txt
( time timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_shard.py -x -q -k "multi_gpu or shard" ) > $O/pytest.log 2>&1; tail -15 $O/pytest.log
python - <<'PY'
import sys; sys.path.insert(0, ".")
from colord_b200 import synth
print(synth.generate_file("/tmp/ns.fastq", "ont", 187500, int(187500 * 8000 / 20.8), 8000, seed=5))
PY
for g in 1 2; do
  if [ $g = 1 ]; then X="--native"; else X="--gpus 2"; fi
  ( time colord_b200/colord-b200 compress-ont $X -v /tmp/ns.fastq /tmp/ns_$g.colord ) > $O/cli_gpus$g.log 2>&1
  grep -E "phase|size|real" $O/cli_gpus$g.log
  ls -l /tmp/ns_$g.colord
done
BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
grep "phase\]" $O/bench_n2.err | tail -16; cut -c1-400 $O/bench_n2.json
