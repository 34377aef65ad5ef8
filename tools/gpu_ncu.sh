#!/bin/bash
# ncu --set full captures of named kernels taken from one eager bs16 forward: args = "regex:skip:count" ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r rx skip cnt <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -s ${skip:-0} -c ${cnt:-1} \
     -o gpurun_out/k_${rx}_${skip:-0} -f python tools/profile_forward.py > gpurun_out/ncu_${rx}_${skip:-0}.log 2>&1
  tail -1 gpurun_out/ncu_${rx}_${skip:-0}.log
done
