# one GPU iteration of the optimisation loop: quick parity, then the two bench workloads (results under gpurun_out/<tag>_*)
tag=${1:-it}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replay or division or tracers_agree" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload xos1 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
tail -3 gpurun_out/${tag}_pytest.log
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
