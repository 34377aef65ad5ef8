#!/bin/bash
# round-2 evidence on one B200: launch list (+ DRAM bytes) of one eager train step, whole-graph counters, in-situ timeline,
# compute-sanitizer memcheck + racecheck
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_train_step_launches.csv python tools/profile_train.py > gpurun_out/r02_profile_train.log 2>&1
python tools/launch_summary.py gpurun_out/r02_train_step_launches.csv > gpurun_out/r02_train_step_launches_summary.txt; head -5 gpurun_out/r02_train_step_launches_summary.txt
python tools/make_traffic.py gpurun_out/r02_train_step_launches.csv gpurun_out/r02_traffic.json
timeout 400 ncu --graph-profiling graph --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,sm__cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --csv --log-file gpurun_out/r02_graph_ncu.csv python tools/r2_graph_ncu.py > gpurun_out/r02_graph_ncu.log 2>&1; tail -1 gpurun_out/r02_graph_ncu.log
timeout 300 python tools/r2_timeline.py r02 > gpurun_out/r02_timeline_summary.txt 2>&1; python tools/timeline_alone.py gpurun_out/r02_timeline.csv x > gpurun_out/r02_timeline_alone.txt 2>&1; head -3 gpurun_out/r02_timeline_summary.txt
timeout 200 python tools/r2_breakdown.py 2>&1 | grep -v -i "warn\|run_backward" > gpurun_out/r02_breakdown.txt; tail -4 gpurun_out/r02_breakdown.txt
timeout 300 python tools/r2_wgrad_sweep.py > gpurun_out/r02_wgrad_sweep.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/r02_sanitizer_racecheck.log
du -sh gpurun_out
