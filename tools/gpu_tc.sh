#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:seg_gemm_tc -s 1 -c 1 -o gpurun_out/seg_tc_r1b python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_seg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 1 -c 1 -o gpurun_out/chain_tc_r1b python tools/prof_pool.py 8 1 1 > gpurun_out/ncu_chain.log 2>&1
ls gpurun_out/*.ncu-rep
