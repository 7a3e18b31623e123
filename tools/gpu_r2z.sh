# round 2: final leak records -- bench line of BASELINE configs[2] with leaks on, raw ncu counters of the phased leak kernel
tag=r2z
mkdir -p gpurun_out
timeout 300 python bench.py --workload ellip --leaks 20000 --no-subs --steps 3 --warmup 3 > gpurun_out/${tag}_ellip_leaks.json 2> gpurun_out/${tag}_ellip_leaks.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_ellip_leaks.json").read().strip().splitlines()[-1]); l=d["leak_run"]
print({k:l[k] for k in l if k not in ("roofline","cpu_baseline","workload")}); print(l.get("cpu_baseline")); print(l.get("roofline"))
PY
export POLYCAP_B200_LEAK_BIG_MAX=64
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio
timeout 200 ncu --metrics $M --clock-control none -k regex:pcb_leak_kernel -c 1 --csv --log-file gpurun_out/${tag}_leak_metrics.csv python tools/leak_probe.py 6000 7 only > gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_ncu.log
