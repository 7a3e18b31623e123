tag=r2l
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_leaks.py tests/test_reference_programs.py::test_reference_leaks_program -q -m gpu --durations=8 ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -22 gpurun_out/${tag}_pytest.log
POLYCAP_B200_PHASES=1 timeout 900 python bench.py --workload ellip --leaks 1000 --no-subs --no-e2e --steps 1 --warmup 1 --no-cpu-baseline --cpu-seconds 15 > gpurun_out/${tag}_ellip_leaks.json 2> gpurun_out/${tag}_ellip_leaks.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2l_ellip_leaks.json").read().strip().splitlines()[0])
print(json.dumps(d.get("leak_run"))[:1500])
PY
grep "phases: leak" gpurun_out/${tag}_ellip_leaks.err | tail -3
