#!/bin/bash
# round-2 session E: new post-processing / KITTI / input-stage tests, predictor-heads fix, quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_postprocess_gpu.py tests/test_kitti_gpu.py -m gpu -q --maxfail=20 --timeout 300 --timeout-method=thread 2>&1 | tail -70 > gpurun_out/pytest_new.log
tail -60 gpurun_out/pytest_new.log
timeout 600 python -m pytest tests/test_gnn_gpu.py -m gpu -q -x -k "golden or reference_graph or T1" --timeout 300 --timeout-method=thread 2>&1 | tail -5
timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_e.log 2>&1; tail -1 gpurun_out/bench_e.log | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_e.log 2>&1
python tools/launch_summary.py gpurun_out/launches_e.csv | head -14
