#!/bin/bash
# session I: ped pooling on tensor cores (chain store mode + pool_last_tc_kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gnn_gpu.py -m gpu -q -x --timeout 300 --timeout-method=thread \
  -k "pool_chain or layers_vs_oracle or predict_matches_golden or every_checkpoint" 2>&1 | tail -30 > gpurun_out/pytest_i.log
tail -30 gpurun_out/pytest_i.log
timeout 300 python bench.py --workload ped_cyl_auto_T3_20k_b8 --steps 20 --no-cpu-baseline > gpurun_out/bench_ped_i.log 2>&1; tail -1 gpurun_out/bench_ped_i.log | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ped_i.csv python bench.py --workload ped_cyl_auto_T3_20k_b8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_ped_i.log 2>&1
python tools/launch_summary.py gpurun_out/launches_ped_i.csv 2>/dev/null | head -16
