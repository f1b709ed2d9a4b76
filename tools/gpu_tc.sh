#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q -k "tc_edge or layers or predict or T1" 2>&1 | tail -40 > gpurun_out/pytest_tc.log
tail -25 gpurun_out/pytest_tc.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1
tail -2 gpurun_out/bench_tc.log
