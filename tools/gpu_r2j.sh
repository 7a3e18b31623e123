tag=r2j
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x --durations=12 ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[0])
print("value %.4e frac %.4f e2e %.4e cold %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["cold_first_call"]))
print("stat_parity", d.get("stat_parity"))
print("cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("sub_records",{}).items(): print(k, {a:v.get(a) for a in ("value","roofline_frac","reflections_per_s","error")})
PY
