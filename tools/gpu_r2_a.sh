#!/bin/bash
# round-2 GPU session: edge-kernel sanity + timing first (bails out if it hangs), parity suite, trace, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 150 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1
rc=$?
tail -3 gpurun_out/prof_edge.log
if [ $rc -ne 0 ]; then echo "prof_edge failed rc=$rc - skipping the rest"; exit 0; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=12 --timeout 420 --timeout-method=thread 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
PG_USE_LAB_LIB=1 PG_TC_TRACE=gpurun_out/trace.txt timeout 200 python tools/prof_edge.py 8 1 1 > gpurun_out/prof_edge_lab.log 2>&1
timeout 100 python tools/trace_seg.py gpurun_out/trace.txt 19 > gpurun_out/trace_summary.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -40 gpurun_out/pytest_gpu.log; cat gpurun_out/trace_summary.txt; tail -1 gpurun_out/bench.log | cut -c1-1800
