# round 2, call B: large-grid parity tests (all), ncu capture of the deferred-weights kernel on xos1
tag=r2b
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity_large.py -q -m gpu ) > gpurun_out/${tag}_pytest_large.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_large.log
tail -15 gpurun_out/${tag}_pytest_large.log
python bench.py --steps 1 --warmup 1 --ids 600000 --workload xos1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain_xos1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_r2b_xos python bench.py --steps 1 --warmup 1 --ids 600000 --workload xos1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_xos1.log 2>&1
ls -la gpurun_out/prof_r2b*
