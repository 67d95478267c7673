#!/bin/bash
# round 2, call Z: the switchable alignment forms against the goldens; compute-sanitizer memcheck (no slab: every buffer its own
# allocation) over the edit-script goldens and one golden of the whole stage 2
set -u
O=gpurun_out/r02z; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_stage2.py -x -q -k "kernel_variants" ) > $O/pytest_variants.log 2>&1; tail -3 $O/pytest_variants.log
( time CLB_SLAB_GB=0 timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_stage2.py -x -q -k "test_edit_scripts_golden or (test_compact_es_bytes_golden and ont_bal and not variants) or (test_compact_es_bytes_golden and hifi and not variants)" ) > $O/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck.log | head -20
