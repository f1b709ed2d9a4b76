#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/prof_edge.log
timeout 300 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_tc.log; tail -4 gpurun_out/pytest_tc.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
