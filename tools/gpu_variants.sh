# bench the library variants under polycap_b200/variants/ (tuning builds made with `make BUILD=... OUT=... EXTRA=...`)
tag=${1:-var}
mkdir -p gpurun_out
for lib in polycap_b200/variants/*.so; do
  v=$(basename $lib .so); v=${v#libpolycap_b200_}
  for w in conical xos1; do
    POLYCAP_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload $w > gpurun_out/${tag}_${v}_$w.json 2> gpurun_out/${tag}_${v}_$w.err
  done
done
for f in gpurun_out/${tag}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%d['value'], round(d['roofline']['frac'],4))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
