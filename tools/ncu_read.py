#!/usr/bin/env python
"""Print selected raw metrics from an .ncu-rep (runs `ncu -i … --page raw --csv`; no GPU needed)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct', 'smsp__average_warps_issue_stalled',
        'sm__throughput.avg.pct', 'sm__inst_executed_pipe_', 'smsp__issue_active.avg.pct', 'sm__pipe_xu', 'sm__pipe_fma_cycles',
        'sm__pipe_alu_cycles', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'launch__grid_size', 'launch__waves', 'dram__throughput.avg.pct', 'lts__t_bytes.sum', 'launch__occupancy_limit',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'l1tex__t_bytes', 'lts__throughput']
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for li, v in enumerate(rows[2:]):
    print('=== launch', li, v[h.index('Kernel Name')][:60] if 'Kernel Name' in h else '')
    for i, k in enumerate(h):
        if any(x in k for x in KEYS + extra):
            if v[i] not in ('0', '', '0.000000'):
                print('  %-90s %s %s' % (k, v[i], rows[1][i]))
