#!/bin/bash
# Kernel-iteration visit: short live fuzz, kernel timings without a profiler, and the forward kernel's executed-instruction / pipe
# counters.  usage (under gpurun): bash tools/gpu_iter.sh <tag> [fuzz_groups] [full]
TAG=${1:-it}; G=${2:-12}
mkdir -p gpurun_out
timeout 600 python tools/gpu_fuzz_live.py $G 991 1 0 > gpurun_out/${TAG}_fuzz.txt 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/${TAG}_fuzz.txt
timeout 300 python tools/prof_run.py 3000 4 > gpurun_out/${TAG}_run.txt 2>&1; tail -2 gpurun_out/${TAG}_run.txt | cut -c1-600
M=smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_elapsed,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:forward_kernel -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_ctr.csv python tools/prof_run.py 3000 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_ctr.csv")) if len(r)>10]
for r in rows[1:]:
    print(r[-3][:70], r[-1])
PY
if [ "$3" == "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_fwd python tools/prof_run.py 3000 2 > gpurun_out/${TAG}_ncu_fwd.log 2>&1; echo "ncu rc=$?"
fi
