#!/bin/bash
# ncu --set full captures of named kernels taken from one eager bs16 TRAIN step (tools/profile_train.py): args = "regex:skip:count" ...
# Only text comes back (gpurun merges at most 64 MiB): the raw-metric CSV of every capture + the selected-metric summary.
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r rx skip cnt <<< "$spec"
  rep=/tmp/r2t_${rx}_${skip:-0}
  timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:$rx -s ${skip:-0} -c ${cnt:-1} \
     -o $rep -f python tools/profile_train.py > gpurun_out/r2t_ncu_${rx}_${skip:-0}.log 2>&1
  tail -1 gpurun_out/r2t_ncu_${rx}_${skip:-0}.log
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/r2t_${rx}_${skip:-0}_raw.csv 2>/dev/null
  python tools/ncu_read.py $rep.ncu-rep dram__bytes sm__cycles_active l1tex__t_sector_hit lts__t_sector_hit smsp__warp_issue_stalled > gpurun_out/r2t_${rx}_${skip:-0}.txt 2>&1
  rm -f $rep.ncu-rep
done
du -sh gpurun_out
