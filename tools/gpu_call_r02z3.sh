#!/bin/bash
# round 2, call Z3: compute-sanitizer memcheck (no slab) over one verbose command-line run in both stream formats (statistics kernels,
# split alignment kernels at a few Mbases, stage 3 native + compat)
set -u
O=gpurun_out/r02z3; mkdir -p $O
( time CLB_SLAB_GB=0 timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_cli.py -x -q -k "verbose_statistics and pbhifi" ) > $O/memcheck_cli.log 2>&1
echo "cli rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_cli.log | sort | uniq -c | head
