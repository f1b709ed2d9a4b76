#!/bin/bash
# session J: TMEM drain microbenchmark + producer lane-map variants of the fused edge kernel
mkdir -p gpurun_out
for w in 4 8; do for v in 0 1 2 3 4 5; do timeout 30 tools/umma_lab drain $w $v; done; done 2>&1 | tee gpurun_out/drain_lab.txt
bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/variants_j.txt
