#!/bin/bash
# round 2, call Y: the -v statistics block against the stock binary (four option sets x both stream formats)
set -u
O=gpurun_out/r02y; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -k "verbose_statistics" ) > $O/pytest.log 2>&1; tail -40 $O/pytest.log
