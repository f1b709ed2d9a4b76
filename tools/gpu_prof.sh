#!/bin/bash
mkdir -p gpurun_out
python tools/prof_edge.py 8 3 1 > gpurun_out/prof_edge.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_gnn_tc -s 1 -c 1 -o gpurun_out/edge_tc python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_edge.log 2>&1
tail -3 gpurun_out/prof_edge.log; tail -3 gpurun_out/ncu_edge.log; ls -la gpurun_out/*.ncu-rep
