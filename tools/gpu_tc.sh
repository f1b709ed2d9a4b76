#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_edge.log
timeout 200 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q -k "tc_edge or full_size or layers_vs" 2>&1 | tail -25 > gpurun_out/pytest_tc.log; tail -4 gpurun_out/pytest_tc.log
PG_TC_TRACE=gpurun_out/trace.txt timeout 60 python tools/prof_edge.py 8 1 1 > gpurun_out/trace_run.log 2>&1
python tools/trace_seg.py gpurun_out/trace.txt 19 2>&1 | head -8 | tee gpurun_out/trace_seg.txt
