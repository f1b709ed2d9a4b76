#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gnn_gpu.py -m gpu -x -q -k "full_size" 2>&1 | tail -15 > gpurun_out/pytest_full.log; tail -5 gpurun_out/pytest_full.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload car_auto_T3_120k > gpurun_out/bench_120k.log 2>&1; tail -1 gpurun_out/bench_120k.log | cut -c1-900
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload ped_cyl_auto_T3_20k_b8 > gpurun_out/bench_ped.log 2>&1; tail -1 gpurun_out/bench_ped.log | cut -c1-900
