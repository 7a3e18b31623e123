tag=r2m
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python tools/multi_device_probe.py conical > gpurun_out/${tag}_probe_conical.json 2> gpurun_out/${tag}_probe_conical.err
python tools/multi_device_probe.py xos1 > gpurun_out/${tag}_probe_xos1.json 2> gpurun_out/${tag}_probe_xos1.err
python bench.py --workload xos1 --no-subs --no-cpu-baseline --steps 3 > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
python - <<'PY'
import json
for w in ("conical","xos1"):
    try:
        d=json.loads(open("gpurun_out/r2m_probe_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "probe speedup %.2f"%d["speedup"], [(r["devices"], "%.4e"%r["photons_per_s"], r["seconds_per_call"]) for r in d["runs"]])
    except Exception as e: print(w, "probe ERR", e, open("gpurun_out/r2m_probe_%s.err"%w).read()[-800:])
d=json.loads(open("gpurun_out/r2m_xos1.json").read().strip().splitlines()[-1]); print("xos1 value %.4e e2e %.4e"%(d["value"], d["e2e"]["value"]), d["e2e"]["seconds_per_call"], d["roofline"]["traffic"])
PY
