# round 2, call E: ncu full captures of the current build (xos1 deferred kernel at bench step size; conical) + launch list
tag=r2e
mkdir -p gpurun_out
cp polycap_b200/libpolycap_b200.so gpurun_out/${tag}_lib.so
A="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-subs"
python bench.py $A --workload xos1 > gpurun_out/${tag}_plain_xos1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_xos python bench.py $A --workload xos1 > gpurun_out/${tag}_ncu_xos1.log 2>&1
python bench.py $A --ids 4000000 > gpurun_out/${tag}_plain_conical.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_conical python bench.py $A --ids 4000000 > gpurun_out/${tag}_ncu_conical.log 2>&1
ls -la gpurun_out/prof_${tag}*
