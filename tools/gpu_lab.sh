#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lab4.log
for t in "rate 208 96 304 0" "rate 208 96 304 3" "rate 208 96 304 1" "rate 160 144 304 3"; do
  timeout 60 ./tools/umma_lab $t >> gpurun_out/lab4.log 2>&1
done
cat gpurun_out/lab4.log
