# round 2: leak kernel -- phase clocks of the current build and one ncu pass of raw counters (instructions, pipes, stalls)
tag=r2r
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_p1.so timeout 300 python tools/leak_probe.py 30000 7 only > gpurun_out/${tag}_p1_30000_7.log 2>&1
grep -E "leak_calc|leak kernel|leak clocks" gpurun_out/${tag}_p1_30000_7.log | cut -c1-400
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_p1.so timeout 300 python tools/leak_probe.py 2000 400 only > gpurun_out/${tag}_p1_2000_400.log 2>&1
grep -E "leak_calc|leak kernel|leak clocks" gpurun_out/${tag}_p1_2000_400.log | cut -c1-400
export POLYCAP_B200_LEAK_BIG_MAX=64
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum,smsp__inst_executed_op_branch.sum,sm__cycles_elapsed.max,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio
timeout 400 ncu --metrics $M --clock-control none -k regex:pcb_leak_kernel -c 1 --csv --log-file gpurun_out/${tag}_leak_metrics.csv python tools/leak_probe.py 6000 7 only > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log; wc -c gpurun_out/${tag}_leak_metrics.csv
