#!/bin/bash
# ncu --set full on stand-alone units; only CSV summaries are kept (reports are too large to bring back)
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warp_latency_issue_stalled_barrier.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"
for unit in "$@"; do
  tag=$(echo $unit | tr ' ' '_')
  python tools/prof_units.py $unit > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed: $unit"; tail -3 gpurun_out/plain_$tag.log; continue; }
  # second repetition only: skip the launches of the first (counts differ per unit; take the tail half)
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_$tag.csv python tools/prof_units.py $unit > gpurun_out/ncu_$tag.log 2>&1
  echo "$unit: $(wc -l < gpurun_out/ncu_$tag.csv) csv lines"
done
