#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], json.dumps(d['roofline_scatter_max']))"
