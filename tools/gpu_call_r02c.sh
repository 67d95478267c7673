#!/bin/bash
# round 2, call C: compat streams on the device — parts against the stock binary's, the command line in both stream formats,
# then size / time of compat against native and the reference on files of growing size
set -u
O=gpurun_out/r02c; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_exact.py -x -q ) > $O/pytest_exact.log 2>&1; tail -15 $O/pytest_exact.log
( time timeout 1200 python -m pytest tests/test_gpu_cli.py -x -q ) > $O/pytest_cli.log 2>&1; tail -15 $O/pytest_cli.log
timeout 900 python tools/ratio_check.py --configs C1,C3 --scale C3:0.0125 --ours-opts=--compat --out $O/ratio_compat_small.json --md $O/ratio_compat_small.md > $O/ratio_compat_small.log 2>&1; cat $O/ratio_compat_small.md
timeout 900 python tools/ratio_check.py --configs C2,C3,NS --scale C2:0.15,C3:0.04,NS:0.04 --ours-opts="--compat -v" --out $O/ratio_compat_mid.json --md $O/ratio_compat_mid.md > $O/ratio_compat_mid.log 2>&1; cat $O/ratio_compat_mid.md
tail -3 $O/ratio_compat_*.log | cut -c1-3000
