#!/bin/bash
# round 2, call B: GPU tests, then the matched-ratio / round-trip table of the command line against the stock reference binary
set -u
O=gpurun_out/r02b; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/smi.txt; nproc >> $O/smi.txt; free -g >> $O/smi.txt; df -h /tmp >> $O/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
timeout 1500 python tools/ratio_check.py --configs C1,C2,C3,C4 --runs 1 --out $O/ratio_c1_c4.json --md $O/ratio_c1_c4.md > $O/ratio_c1_c4.log 2>&1
cat $O/ratio_c1_c4.md
timeout 1200 python tools/ratio_check.py --configs NS --runs 3 --t1 --out $O/ratio_ns.json --md $O/ratio_ns.md > $O/ratio_ns.log 2>&1
cat $O/ratio_ns.md
tail -5 $O/ratio_*.log | cut -c1-1500
