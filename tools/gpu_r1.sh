#!/bin/bash
# One GPU session: tests, smoke, bench (+ reference arm), ncu launch list, ncu full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt; nproc >> gpurun_out/cpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_gemm_tc -s 3 -c 1 -o gpurun_out/edge_tc_r1 python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_edge.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_gemm_tc -s 1 -c 1 -o gpurun_out/pool_tc_r1 python tools/prof_pool.py 8 1 1 > gpurun_out/ncu_pool.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-1500; tail -1 gpurun_out/bench_ref.log | cut -c1-600
ls -la gpurun_out/
