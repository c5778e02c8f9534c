# session 5, call b: pure-register NTT round microbenchmark (what the instruction stream reaches without memory phases)
cd tools/ubench
./ntt_math | tee ../../gpurun_out/s5b_ntt_math.log
M=sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
timeout 600 ncu --metrics $M --clock-control none --csv --log-file ../../gpurun_out/s5b_ntt_math_ncu.csv ./ntt_math > /dev/null 2>&1
tail -3 ../../gpurun_out/s5b_ntt_math_ncu.csv | cut -c1-300
