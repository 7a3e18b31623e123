# round 2, call C: full GPU suite (incl. multi-device shares, chunked downloads), benches with e2e
tag=r2c
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
POLYCAP_B200_PHASES=1 $B --workload xos1 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
$B --workload ellip > gpurun_out/${tag}_ellip.json 2> gpurun_out/${tag}_ellip.err
grep phases gpurun_out/${tag}_xos1.err | tail -14
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4), "refl/s %.3e"%d['reflections_per_s'], "e2e %.4e"%d['e2e']['value'], d['e2e']['seconds_per_call'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
