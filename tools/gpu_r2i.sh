tag=r2i
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
python -m pytest tests/test_gpu_parity_large.py tests/test_gpu_leaks.py -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
mv gpurun_out/parity_margins.txt gpurun_out/${tag}_margins_default.txt
POLYCAP_B200_LEAK_DEPTH1=2 python -m pytest tests/test_gpu_leaks.py -q -m gpu > gpurun_out/${tag}_pytest_leak_retry.log 2>&1; tail -2 gpurun_out/${tag}_pytest_leak_retry.log
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_o2.so python -m pytest tests/test_gpu_parity_large.py -q -m gpu -k "per_photon or one_million" > gpurun_out/${tag}_pytest_o2.log 2>&1; tail -2 gpurun_out/${tag}_pytest_o2.log
mv gpurun_out/parity_margins.txt gpurun_out/${tag}_margins_o2.txt
$B --workload xos1 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
$B > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_o2.so $B --workload xos1 > gpurun_out/${tag}_xos1_o2.json 2> gpurun_out/${tag}_xos1_o2.err
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_o2.so $B > gpurun_out/${tag}_conical_o2.json 2> gpurun_out/${tag}_conical_o2.err
POLYCAP_B200_PHASES=1 timeout 600 python bench.py --workload ellip --leaks 1000 --no-subs --no-e2e --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ellip_leaks.json 2> gpurun_out/${tag}_ellip_leaks.err
cat gpurun_out/${tag}_margins_default.txt; echo ---o2; cat gpurun_out/${tag}_margins_o2.txt
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4), "refl/s %.3e"%d['reflections_per_s'], d['clocks']['sm_mhz'], d['clocks']['reasons'], json.dumps(d.get('leak_run'))[:900] if d.get('leak_run') else '')
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
grep phases gpurun_out/${tag}_ellip_leaks.err | tail -5
