#!/bin/bash
# ncu --set full captures of named kernels taken from one eager bs16 TRAIN step (tools/profile_train.py): args = "regex:skip:count" ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r rx skip cnt <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -s ${skip:-0} -c ${cnt:-1} \
     -o gpurun_out/r2t_${rx}_${skip:-0} -f python tools/profile_train.py > gpurun_out/r2t_ncu_${rx}_${skip:-0}.log 2>&1
  tail -1 gpurun_out/r2t_ncu_${rx}_${skip:-0}.log
done
