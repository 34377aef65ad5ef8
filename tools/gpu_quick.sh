#!/bin/bash
# quick GPU visit: a subset of parity tests (pattern $1) under a short timeout, then a short bench
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -25 gpurun_out/pytest_quick.log
if [ -n "$2" ]; then
  timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
  cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
fi
