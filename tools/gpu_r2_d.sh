#!/bin/bash
# round-2 session D: parity suite, smoke, bench (+ reference arm), other workloads, ncu launch list + full capture
mkdir -p gpurun_out
timeout 150 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1 || { tail -5 gpurun_out/prof_edge.log; echo "prof_edge failed"; exit 0; }
tail -1 gpurun_out/prof_edge.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=8 --timeout 420 --timeout-method=thread 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 300 python bench.py --workload car_auto_T3_120k --steps 10 --no-cpu-baseline > gpurun_out/bench_120k.log 2>&1; tail -1 gpurun_out/bench_120k.log | cut -c1-600
timeout 300 python bench.py --workload ped_cyl_auto_T3_20k_b8 --steps 10 --no-cpu-baseline > gpurun_out/bench_ped.log 2>&1; tail -1 gpurun_out/bench_ped.log | cut -c1-600
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:seg_gemm_tc -s 1 -c 1 -o gpurun_out/seg_tc python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_seg.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
