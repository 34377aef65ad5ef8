#!/bin/bash
# round-end check on one B200: GPU tests, smoke, both bench arms (forward headline + train_step leg), train-mode bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --mode train --impl reference --steps 1 --warmup 1 > gpurun_out/bench_train_ref.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -c 1600 gpurun_out/bench.json; echo; cat gpurun_out/bench_train.json | head -c 600; echo; cat gpurun_out/bench_train_ref.json | head -c 400; echo; tail -3 gpurun_out/bench.err
