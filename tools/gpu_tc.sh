#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_edge.log
PG_TC_ROWMAJOR=1 timeout 90 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge_rowmajor.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_edge_rowmajor.log
timeout 150 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q -k "tc_edge_kernel or fused_equals" 2>&1 | tail -15 > gpurun_out/pytest_tc_a.log; tail -4 gpurun_out/pytest_tc_a.log
timeout 90 python tools/prof_pool.py 8 5 1 > gpurun_out/prof_pool.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_pool.log
timeout 240 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_tc.log; tail -6 gpurun_out/pytest_tc.log
