#!/bin/bash
# session N: full parity suite + bench lines (car, ped, 120k) + launch list after the seg / ped / multi-scale work
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 --durations=5 --timeout 420 --timeout-method=thread 2>&1 | tail -14 > gpurun_out/pytest_n.log
tail -14 gpurun_out/pytest_n.log
timeout 600 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_n.log 2>&1; tail -1 gpurun_out/bench_n.log | cut -c1-900
timeout 300 python bench.py --workload ped_cyl_auto_T3_20k_b8 --steps 30 --no-cpu-baseline > gpurun_out/bench_ped_n.log 2>&1; tail -1 gpurun_out/bench_ped_n.log | cut -c1-300
timeout 300 python bench.py --workload car_auto_T3_120k --steps 30 --no-cpu-baseline > gpurun_out/bench_120k_n.log 2>&1; tail -1 gpurun_out/bench_120k_n.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_n.log 2>&1
python tools/launch_summary.py gpurun_out/launches_n.csv 2>/dev/null | head -14
