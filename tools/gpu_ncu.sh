# one ncu --set full capture of the trace kernel per workload (after the same command ran clean without ncu)
tag=${1:-p}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --ids 2000000 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_r1_${tag} python bench.py --steps 1 --warmup 1 --ids 2000000 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu.log 2>&1
if [ "$2" = "xos1" ]; then
python bench.py --steps 1 --warmup 1 --ids 600000 --workload xos1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain_xos1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_r1_${tag}_xos python bench.py --steps 1 --warmup 1 --ids 600000 --workload xos1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_xos1.log 2>&1
fi
ls -la gpurun_out/prof_r1_${tag}*
