tag=r2g
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
for v in t512eu2 t512eu3 t512eu4 t448eu4 t384eu4; do
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so $B --workload xos1 > gpurun_out/${tag}_xos1_$v.json 2> gpurun_out/${tag}_xos1_$v.err
done
for v in t512eu3 t448eu4; do
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so $B --workload ellip > gpurun_out/${tag}_ellip_$v.json 2> gpurun_out/${tag}_ellip_$v.err
done
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4), "refl/s %.3e"%d['reflections_per_s'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
