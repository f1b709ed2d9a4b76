#!/bin/bash
# check + time + trace the fused edge layer for every experiment build in lab/*.so
mkdir -p gpurun_out
for so in lab/*.so; do
  v=$(basename $so .so)
  PG_LIB_VARIANT=$v timeout 200 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_$v.log 2>&1
  echo "== $v: $(tail -2 gpurun_out/prof_$v.log | tr '\n' ' ')"
  PG_LIB_VARIANT=$v PG_TC_TRACE=gpurun_out/trace_$v.txt timeout 150 python tools/prof_edge.py 8 1 1 > gpurun_out/proft_$v.log 2>&1
  timeout 60 python tools/trace_seg.py gpurun_out/trace_$v.txt 19 2>&1 | head -14
done
