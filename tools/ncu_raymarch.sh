#!/bin/bash
# One ncu capture of the shipped raymarch kernel (<false> variant) inside bench.py's cfg3 workload; selected counters as CSV.
# usage: tools/ncu_raymarch.sh <tag>     -> gpurun_out/<tag>_raymarch_ncu.csv
TAG=${1:-cur}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
timeout 600 ncu --clock-control none --metrics $M -k regex:raymarch_kernel -s 12 -c 1 --csv --log-file gpurun_out/${TAG}_raymarch_ncu.csv \
  python bench.py --no-mesh --no-cpu --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_raymarch_ncu.csv")) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print("%-95s %-10s %s" % (d.get("Metric Name"), d.get("Metric Unit"), d.get("Metric Value")))
print(rows[1][h.index("Kernel Name")][:60] if len(rows)>1 else "no rows")
PY
