#!/bin/bash
# round 2, call M: k_align as a forward and a backward kernel (the backward one at 24 warps per SM, its walk prefetching into L2):
# stage-2 parity + shards + command line, phase split, short bench, synchronising trace of every stage at 6 Gbases, ncu of a bulk forward launch
set -u
O=gpurun_out/r02m; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_shard.py -x -q ) > $O/pytest_stage2.log 2>&1; tail -5 $O/pytest_stage2.log
if grep -q "failed\|error" $O/pytest_stage2.log; then grep -E "Error|assert|FAILED" $O/pytest_stage2.log | head -20; exit 1; fi
CLB_LIBRARY=$PWD/colord_b200/libcolord_b200_phases.so timeout 600 python bench.py --gbases 6 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/phases.json 2> $O/phases.err
grep "align phases" $O/phases.err | tail -6
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02m/bench_short.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
CLB_S2_TRACE=1 BENCH_PHASES=1 timeout 600 python bench.py --gbases 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/trace.json 2> $O/trace.err
grep -E "s3q|s3d|\[phase\]" $O/trace.err | tail -40
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base function -k k_align --launch-skip 176 --launch-count 3 -o $O/k_align_fwd32 -f \
  python bench.py --gbases 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_bench.json 2> $O/ncu_bench.err
ncu -i $O/k_align_fwd32.ncu-rep --page raw --csv > $O/k_align_fwd32_raw.csv 2>/dev/null
ncu -i $O/k_align_fwd32.ncu-rep --page source --csv --print-source cuda,sass > $O/k_align_fwd32_source.csv 2>/dev/null
ls -l $O; [ $(stat -c %s $O/k_align_fwd32.ncu-rep) -gt 40000000 ] && rm $O/k_align_fwd32.ncu-rep; true
