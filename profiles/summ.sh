#!/bin/bash
# usage (here, no GPU): profiles/summ.sh <report.ncu-rep> <kernel-substring>
ncu -i $1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for i,name in enumerate(h):
    if name in want: print(f'{name} = {v[i]} {rows[1][i]}')
for i,name in enumerate(h):
    if name.startswith('smsp__average_warps_issue_stalled') and name.endswith('per_issue_active.ratio') and float(v[i] or 0) > 0.3: print(f'{name[34:-24]} = {v[i]}')
"
rm -rf /tmp/vpt_cubin && mkdir -p /tmp/vpt_cubin && (cd /tmp/vpt_cubin && cuobjdump -xelf all /root/repo/vaporetto_b200/libvaporetto_b200.so > /dev/null 2>&1)
python /root/repo/profiles/line_profile.py $1 $2 /tmp/vpt_cubin/$(ls /tmp/vpt_cubin/*${4:-kernels}*.cubin | head -1 | xargs basename) ${3:-30}
