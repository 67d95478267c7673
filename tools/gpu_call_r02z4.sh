#!/bin/bash
# round 2, call Z4 (2 GPUs): the multi-GPU command-line tests with the statistics block of the shards against the one-GPU block
set -u
O=gpurun_out/r02z4; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -k "multi_gpu" ) > $O/pytest_mgpu.log 2>&1; tail -15 $O/pytest_mgpu.log
