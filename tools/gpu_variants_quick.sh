#!/bin/bash
# time the fused edge layer for every experiment build in lab/*.so (no trace), three repetitions each
mkdir -p gpurun_out
for rep in 1 2; do
for so in lab/*.so; do
  v=$(basename $so .so)
  PG_LIB_VARIANT=$v timeout 200 python tools/prof_edge.py 8 8 1 > gpurun_out/prof_$v.log 2>&1
  echo "== $v: $(tail -1 gpurun_out/prof_$v.log)"
done
done
