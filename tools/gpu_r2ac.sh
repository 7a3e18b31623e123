# round 2: T2_SCAN_AFTER variants of the task-queue tracer (segment steps a photon takes inside the task that reflected it)
mkdir -p gpurun_out
A="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
for v in s6 s2; do
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so timeout 40 python bench.py $A > gpurun_out/r2ac_$v.json 2>/dev/null
  python - gpurun_out/r2ac_$v.json $v <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "%.4e photons/s"%d["value"], "frac %.4f"%d["roofline"]["frac"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
done
