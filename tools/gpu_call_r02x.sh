#!/bin/bash
# round 2, call X: BASELINE configurations 3 and 4 at full size through the reworked path (one step each), the command line on a 3 GB
# north-star slice (file -> archive, round trip), and the stock binary on the same file
set -u
O=gpurun_out/r02x; mkdir -p $O
for cfg in C3 C4; do
  timeout 600 python bench.py --config $cfg --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  python - $cfg <<'PY'
import json, sys
try:
    l = json.loads([x for x in open(f"gpurun_out/r02x/bench_{sys.argv[1]}.json") if x.startswith("{")][-1])
    print(sys.argv[1], round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", sys.argv[1], e)
PY
  tail -2 $O/bench_$cfg.err
done
python - <<'PY'
import sys; sys.path.insert(0, ".")
from colord_b200 import synth
print(synth.generate_file("/tmp/ns.fastq", "ont", 187500, int(187500 * 8000 / 20.8), 8000, seed=5))
PY
ls -l /tmp/ns.fastq
( time timeout 300 colord_b200/colord-b200 compress-ont --native -v /tmp/ns.fastq /tmp/ns.colord ) > $O/cli_native.log 2>&1; grep -E "phase|size|real|rror" $O/cli_native.log; ls -l /tmp/ns.colord
( time timeout 300 colord_b200/colord-b200 decompress /tmp/ns.colord /tmp/ns.back ) > $O/cli_decompress.log 2>&1; grep -E "real|rror" $O/cli_decompress.log
python - <<'PY'
# bases and headers of the round trip equal the input (qualities are the 4-avg transform)
a = open("/tmp/ns.fastq", "rb"); b = open("/tmp/ns.back", "rb"); n = bad = 0
while True:
    la, lb = a.readline(), b.readline()
    if not la or not lb: break
    if n % 4 in (0, 1) and la != lb: bad += 1
    n += 1
print("round trip:", n // 4, "records,", bad, "differing header/base lines,", "tail ok" if not a.readline() and not b.readline() else "LENGTH DIFFERS")
PY
( time timeout 600 oracle/_ref/colord compress-ont -t $(nproc) /tmp/ns.fastq /tmp/ns.ref.colord ) > $O/cli_reference.log 2>&1; grep -E "real" $O/cli_reference.log; ls -l /tmp/ns.ref.colord
