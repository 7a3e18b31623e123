# round 2, call A: large-grid parity tests, deferred-weights kernel on/off, variants -> gpurun_out/r2a_*
tag=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt; nproc >> gpurun_out/${tag}_gpu.txt
( time python -m pytest tests/test_gpu_parity_large.py -x -q -m gpu ) > gpurun_out/${tag}_pytest_large.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_large.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replay or tracers_agree or per_photon or exact_checks" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
$B --workload xos1 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
POLYCAP_B200_DEFER_MIN=0 $B --workload xos1 > gpurun_out/${tag}_xos1_nodefer.json 2> gpurun_out/${tag}_xos1_nodefer.err
POLYCAP_B200_DEFER_HK=32 $B --workload xos1 > gpurun_out/${tag}_xos1_hk32.json 2> gpurun_out/${tag}_xos1_hk32.err
$B --workload ellip > gpurun_out/${tag}_ellip.json 2> gpurun_out/${tag}_ellip.err
for v in eu3 eu4; do
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so $B --workload xos1 > gpurun_out/${tag}_xos1_$v.json 2> gpurun_out/${tag}_xos1_$v.err
done
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_bo512.so $B > gpurun_out/${tag}_conical_bo512.json 2> gpurun_out/${tag}_conical_bo512.err
tail -4 gpurun_out/${tag}_pytest_large.log; tail -2 gpurun_out/${tag}_pytest.log
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4), "refl/s %.3e"%d['reflections_per_s'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
