#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_gemm_tc -s 3 -c 1 -o gpurun_out/edge_tc3 python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_edge.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_edge.log
