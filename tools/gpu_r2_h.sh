#!/bin/bash
# session H: single-traversal graph build + 4-lane dense producers: full parity suite, bench, launch list
mkdir -p gpurun_out
timeout 150 python tools/prof_edge.py 8 5 1 > gpurun_out/prof_edge.log 2>&1 || { tail -5 gpurun_out/prof_edge.log; echo "prof_edge failed"; exit 0; }
tail -2 gpurun_out/prof_edge.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 --durations=6 --timeout 420 --timeout-method=thread 2>&1 | tail -50 > gpurun_out/pytest_h.log
tail -40 gpurun_out/pytest_h.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_h.log 2>&1; tail -1 gpurun_out/bench_h.log | cut -c1-1700
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_h.log 2>&1
python tools/launch_summary.py gpurun_out/launches_h.csv 2>/dev/null | head -24
