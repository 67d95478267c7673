#!/bin/bash
# round 2, call K: the whole GPU suite with the QB02 quality container, a short bench, the phase split of k_align (profiling build
# with cycle counters), and one ncu --set full capture with source of a bulk k_align<32> launch
set -u
O=gpurun_out/r02k; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
CLB_LIBRARY=$PWD/colord_b200/libcolord_b200_phases.so timeout 600 python bench.py --gbases 6 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/phases.json 2> $O/phases.err
grep "align phases" $O/phases.err | tail -6
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err
python - <<'PY'
import json
try:
    l = json.loads([x for x in open("gpurun_out/r02k/bench_short.json") if x.startswith("{")][-1])
    print(round(l["value"]), "MB/s", round(l["ms_per_step"]), "ms", {k: round(v) for k, v in l["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base function -k k_align --launch-skip 90 --launch-count 1 -o $O/k_align32_bulk -f \
  python bench.py --gbases 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_bench.json 2> $O/ncu_bench.err
ls -l $O/*.ncu-rep
ncu -i $O/k_align32_bulk.ncu-rep --page raw --csv > $O/k_align32_bulk_raw.csv 2>/dev/null
ncu -i $O/k_align32_bulk.ncu-rep --page source --csv --print-source cuda,sass > $O/k_align32_bulk_source.csv 2>/dev/null || ncu -i $O/k_align32_bulk.ncu-rep --page source --csv > $O/k_align32_bulk_source.csv 2>/dev/null
ls -l $O; [ $(stat -c %s $O/k_align32_bulk.ncu-rep) -gt 40000000 ] && rm $O/k_align32_bulk.ncu-rep
