#!/bin/bash
# round 2, call Z2: compute-sanitizer memcheck (no slab) over stage 1 (one-pass count export), all stage-2 goldens, and one verbose
# command-line run (statistics kernels, both stream formats)
set -u
O=gpurun_out/r02z2; mkdir -p $O
( time CLB_SLAB_GB=0 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_stage1.py -x -q ) > $O/memcheck_stage1.log 2>&1
echo "stage1 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_stage1.log | head
( time CLB_SLAB_GB=0 timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_stage2.py -x -q -k "golden and not variants" ) > $O/memcheck_stage2.log 2>&1
echo "stage2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_stage2.log | head
( time CLB_SLAB_GB=0 timeout 500 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_cli.py -x -q -k "verbose_statistics and ratio" ) > $O/memcheck_cli.log 2>&1
echo "cli rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_cli.log | sort | uniq -c | head
