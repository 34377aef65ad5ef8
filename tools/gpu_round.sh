#!/bin/bash
# One GPU visit: parity tests, bench, launch list. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/ncu_fwd.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -2 gpurun_out/ncu_fwd.log
