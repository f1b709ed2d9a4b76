#!/bin/bash
# session G: split-drain seg kernel: numeric check + timing + trace, then the edge/GNN/graph tests
mkdir -p gpurun_out
bash tools/gpu_variants.sh
timeout 150 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1 || { tail -5 gpurun_out/prof_edge.log; echo "prof_edge failed"; exit 0; }
tail -2 gpurun_out/prof_edge.log
timeout 1500 python -m pytest tests/test_gnn_gpu.py tests/test_graph_gpu.py -m gpu -q --maxfail=10 --durations=5 --timeout 420 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/pytest_g.log
tail -30 gpurun_out/pytest_g.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_g.log 2>&1; tail -1 gpurun_out/bench_g.log | cut -c1-2400
