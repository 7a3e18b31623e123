# ncu launch list + one full capture per workload of the current build (after the same commands ran clean without ncu)
tag=${1:-prof}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --ids 2000000 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --ids 2000000 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
bash tools/gpu_ncu.sh ${tag} xos1
