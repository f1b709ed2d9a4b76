#!/bin/bash
bash tools/gpu_variants_quick.sh 2>&1 | tail -5
PG_LIB_VARIANT=sa6 PG_TC_TRACE=gpurun_out/trace_sa6.txt timeout 150 python tools/prof_edge.py 8 1 1 > gpurun_out/proft_sa6.log 2>&1
python tools/trace_seg.py gpurun_out/trace_sa6.txt 19 | head -14
