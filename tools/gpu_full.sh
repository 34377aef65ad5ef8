#!/bin/bash
# Full GPU visit: parity tests, smoke, bench (both arms), module timings, launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q -rP > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 200 python tools/modbench.py > gpurun_out/modbench.txt 2>&1
timeout 200 python tools/microbench.py > gpurun_out/microbench.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/ncu_fwd.log 2>&1
bash tools/gpu_ncu.sh gemm_tc:0:2 gemm_tc:40:2 dwln_kernel:0:1 mb_fused16:0:1 ea16_packT:0:1 > gpurun_out/ncu_kernels.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_tc -s 1 -c 1 -o gpurun_out/flash16_full -f python tools/ncu_one.py flash16 > gpurun_out/ncu_flash16.log 2>&1
grep -E "passed|failed|rc=" gpurun_out/pytest.log | tail -3; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
