#!/bin/bash
mkdir -p gpurun_out
PG_TC_TRACE=gpurun_out/trace.txt python tools/prof_edge.py 8 1 1 > gpurun_out/trace_run.log 2>&1
python tools/trace_summary.py gpurun_out/trace.txt 19 > gpurun_out/trace_summary.txt 2>&1
cat gpurun_out/trace_summary.txt; tail -2 gpurun_out/trace_run.log
