#!/bin/bash
# Final session of the round: full parity suite, smoke, bench lines (car default incl. CPU baseline, ped, 120k, reference
# arm), launch list and one full ncu capture of each fused kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=5 --timeout 420 --timeout-method=thread 2>&1 | tail -14 > gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-600
timeout 300 python bench.py --workload ped_cyl_auto_T3_20k_b8 --steps 30 --no-cpu-baseline > gpurun_out/bench_ped_final.log 2>&1; tail -1 gpurun_out/bench_ped_final.log | cut -c80-200
timeout 300 python bench.py --workload car_auto_T3_120k --steps 30 --no-cpu-baseline > gpurun_out/bench_120k_final.log 2>&1; tail -1 gpurun_out/bench_120k_final.log | cut -c80-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; tail -1 gpurun_out/bench_ref_final.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_final.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:seg_gemm_tc -s 1 -c 1 -f -o gpurun_out/seg_tc python tools/prof_edge.py 8 1 1 > gpurun_out/ncu_seg.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 1 -c 1 -f -o gpurun_out/chain_tc python tools/prof_pool.py 8 1 1 > gpurun_out/ncu_chain.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pool_last -s 1 -c 1 -f -o gpurun_out/pool_last_tc python tools/prof_pool.py 8 1 1 ped_cyl_auto_T3_trainval > gpurun_out/ncu_pool_last.log 2>&1
ls -la gpurun_out/*.ncu-rep
