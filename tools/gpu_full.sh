#!/bin/bash
# Full GPU visit: parity tests, bench (both arms), launch list, microbench, ncu --set full of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 200 python tools/microbench.py > gpurun_out/microbench.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/ncu_fwd.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash -c 2 -o gpurun_out/flash_full -f \
   python tools/ncu_one.py flash > gpurun_out/ncu_flash.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 2 -o gpurun_out/gemm_full -f \
   python tools/ncu_one.py linear 50176 256 64 > gpurun_out/ncu_gemm.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err; tail -2 gpurun_out/ncu_fwd.log
