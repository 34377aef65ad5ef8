#!/bin/bash
# quick GPU visit: parity tests (pattern $1, "all" = everything) then optional microbench ($2, "-" to skip) / bench ($3)
mkdir -p gpurun_out
K="$1"; if [ "$K" = "all" ]; then K=""; fi
timeout 600 python -m pytest tests -m gpu -x -q -rP -k "$K" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
grep -E "logits max-abs|passed|failed|Error|error|rc=" gpurun_out/pytest_quick.log | tail -30
if [ -n "$2" ] && [ "$2" != "-" ]; then timeout 200 python tools/microbench.py $2 > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt; fi
if [ -n "$3" ]; then
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
  cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/ncu_fwd.log 2>&1
  python tools/launch_summary.py gpurun_out/launches.csv | head -40
fi
