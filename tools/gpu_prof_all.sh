#!/bin/bash
# Profiles that back bench.py's numbers: launch list of the bench command + one full capture of each fused kernel.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:seg_gemm_tc -s 1 -c 1 -o gpurun_out/seg_tc python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_seg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 1 -c 1 -o gpurun_out/chain_tc python tools/prof_pool.py 8 1 1 > gpurun_out/ncu_chain.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
