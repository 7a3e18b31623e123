# round-end measurement set: both bench arms, the other workloads, ncu launch list and full captures -> gpurun_out/<tag>_*
tag=${1:-final}
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
python bench.py > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
python bench.py --workload xos1 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
python bench.py --workload ellip > gpurun_out/${tag}_ellip.json 2> gpurun_out/${tag}_ellip.err
python bench.py --steps 1 --warmup 3 --ids 2000000 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --ids 2000000 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
bash tools/gpu_ncu.sh ${tag} xos1
for f in gpurun_out/${tag}_ref.json gpurun_out/${tag}_conical.json gpurun_out/${tag}_xos1.json gpurun_out/${tag}_ellip.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
