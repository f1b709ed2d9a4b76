#!/bin/bash
mkdir -p gpurun_out
for t in "1cta 0" "1cta 1" "2cta 0" "2cta 1" "redux"; do
  echo "== umma_lab $t" >> gpurun_out/lab.log
  timeout 60 ./tools/umma_lab $t >> gpurun_out/lab.log 2>&1
  echo "rc=$?" >> gpurun_out/lab.log
done
cat gpurun_out/lab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --frames-per-step 2 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_fp32.csv
