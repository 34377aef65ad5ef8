#!/bin/bash
# round-end check on one B200: GPU tests, smoke, both bench arms (the fwd+bwd headline with its forward leg), 256-geometry train bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --mode forward --no-gpu-baseline > gpurun_out/bench_forward.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --size 256 --classes 2 --in-ch 3 --no-cpu --no-forward --no-gpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_train_256.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; head -c 700 gpurun_out/bench.json; echo; head -c 500 gpurun_out/bench_ref.json; echo; head -c 400 gpurun_out/bench_forward.json; echo; head -c 400 gpurun_out/bench_train_256.json; echo; tail -3 gpurun_out/bench.err
