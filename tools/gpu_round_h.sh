#!/bin/bash
# source-level ncu look at the bandwidth kernels (pool / lrn / pack): top stall lines + memory-system summary, exported on the box
mkdir -p gpurun_out
T=${1:-r01m}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'pool_plane|lrn_kernel' -c 6 -o /tmp/${T}_pw -f python tools/pw_bench.py > gpurun_out/${T}_ncu_pw.log 2>&1
ncu -i /tmp/${T}_pw.ncu-rep --page raw --csv > /tmp/${T}_pw_raw.csv 2>/dev/null
python - <<PY
import csv
rows=[r for r in csv.reader(open('/tmp/${T}_pw_raw.csv')) if r]
hdr=next(r for r in rows if 'Kernel Name' in r); hi=rows.index(hdr); units=rows[hi+1]; ix={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__average_warp_latency_issue_stalled_long_scoreboard_not_issued.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor']
for r in rows[hi+2:]:
    if len(r)!=len(hdr): continue
    print(r[ix['Kernel Name']][:46], r[ix['launch__grid_size']] if 'launch__grid_size' in ix else '')
    for k in want:
        if k in ix and r[ix[k]]!='': print('    %-90s %s %s'%(k,r[ix[k]],units[ix[k]]))
PY
for id in 0 3; do
ncu -i /tmp/${T}_pw.ncu-rep --page source --csv --kernel-id :::$((id+1)) > /tmp/${T}_src_$id.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/${T}_src_$id.csv')))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)>ix['# Samples'] and r[ix['# Samples']].isdigit()]
tot=sum(int(r[ix['# Samples']]) for r in data)
print(rows[0][1][:60], 'total samples',tot,'instrs',len(data))
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print('  stall totals:', {k:v for k,v in sorted(agg.items(), key=lambda t:-t[1])[:6]})
for i,r in sorted(sorted(enumerate(data), key=lambda t:-int(t[1][ix['# Samples']]))[:12]):
    print('   ',i, r[ix['# Samples']], r[ix['Source']][:90])
PY
done
