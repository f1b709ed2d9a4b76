#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_edge.log
PG_TC_ROUNDROBIN=1 timeout 60 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge_rr.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_edge_rr.log
timeout 200 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q -k "tc_edge or full_size or layers_vs or golden" 2>&1 | tail -25 > gpurun_out/pytest_tc.log; tail -4 gpurun_out/pytest_tc.log
