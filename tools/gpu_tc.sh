#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_tc.log
tail -4 gpurun_out/pytest_tc.log
python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1; tail -1 gpurun_out/prof_edge.log
python tools/prof_pool.py 8 5 1 > gpurun_out/prof_pool.log 2>&1; tail -1 gpurun_out/prof_pool.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1
tail -1 gpurun_out/bench_tc.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_gemm_tc -s 3 -c 1 -o gpurun_out/edge_tc4 python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_edge.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_gemm_tc -s 1 -c 1 -o gpurun_out/pool_tc4 python tools/prof_pool.py 8 1 1 > gpurun_out/ncu_pool.log 2>&1
ls gpurun_out/*.ncu-rep
