tag=r2h
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
python -m pytest tests/test_gpu_parity_large.py -q -m gpu -x -k "deferred or xos1_full or ellipsoidal_400" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
$B --workload xos1 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
$B --workload ellip > gpurun_out/${tag}_ellip.json 2> gpurun_out/${tag}_ellip.err
$B --workload dubfoc > gpurun_out/${tag}_dubfoc.json 2> gpurun_out/${tag}_dubfoc.err
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_hldg.so $B --workload xos1 > gpurun_out/${tag}_xos1_hldg.json 2> gpurun_out/${tag}_xos1_hldg.err
$B > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4), "refl/s %.3e"%d['reflections_per_s'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
