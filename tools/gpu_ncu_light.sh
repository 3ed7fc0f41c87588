#!/bin/bash
# light ncu pass (a few metrics, one launch) for several library variants: tools/gpu_ncu_light.sh "" _tl
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,smsp__thread_inst_executed_per_inst_executed.ratio
for s in "$@"; do
  echo "== variant '$s'"
  VRT_LIB_SUFFIX=$s ncu --metrics $M --clock-control none -k regex:k_trace_camera -s 2 -c 1 python tools/ncu_target.py 3 2>&1 | grep -E "k_trace|duration|hit_rate|sectors|issue_active|inst_executed|scoreboard|latency" 
done 2>&1 | tee gpurun_out/ncu_light.log
