#!/bin/bash
# round-2 session A: parity suite (new full-size tests), edge-kernel timing after the warp-priority change, trace, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -45 > gpurun_out/pytest_gpu.log
timeout 200 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1
PG_USE_LAB_LIB=1 PG_TC_TRACE=gpurun_out/trace.txt timeout 200 python tools/prof_edge.py 8 1 1 > gpurun_out/prof_edge_lab.log 2>&1
timeout 100 python tools/trace_seg.py gpurun_out/trace.txt 19 > gpurun_out/trace_summary.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/prof_edge.log | tail -2; cat gpurun_out/trace_summary.txt; tail -1 gpurun_out/bench.log | cut -c1-1800
