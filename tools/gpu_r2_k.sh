#!/bin/bash
# session K: faster D1 drain (runs precomputed before the wait, 8-column groups on the boundary path, ids one tile ahead)
mkdir -p gpurun_out
bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/variants_k.txt
timeout 900 python -m pytest tests/test_gnn_gpu.py -m gpu -q -x --timeout 300 --timeout-method=thread \
  -k "tails or layers_vs_oracle or predict_matches_golden or fused_equals" 2>&1 | tail -8 | tee gpurun_out/pytest_k.log
