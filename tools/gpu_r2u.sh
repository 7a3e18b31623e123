# round 2: trace2 claim retries / poll back-off variants; leak e2e after the host-side record path change
tag=r2u
mkdir -p gpurun_out
A="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
for v in default t6 b1k t6b1k b16k; do
  if [ $v = default ]; then unset POLYCAP_B200_LIB; else export POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so; fi
  for w in conical; do
    timeout 200 python bench.py $A --workload $w > gpurun_out/${tag}_${v}_$w.json 2> gpurun_out/${tag}_${v}_$w.err
    python - gpurun_out/${tag}_${v}_$w.json $v $w <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["counters"]
    print(sys.argv[2], sys.argv[3], "%.4e photons/s"%d["value"], "frac %.4f"%d["roofline"]["frac"], "batches %.3e lost %.3e empty %.3e"%(c["q_batches"], c["q_lost_races"], c["q_empty_polls"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
done
unset POLYCAP_B200_LIB
export POLYCAP_B200_PHASES=1
for a in "20000 400"; do
  f=gpurun_out/${tag}_leak_$(echo $a | tr ' ' '_').log
  timeout 300 python tools/leak_probe.py $a only > $f 2>&1
  echo "== leak $a"; grep -E "leak_calc|leak kernel|phases: wave|phases: optic" $f | cut -c1-330
done
timeout 300 python -m pytest tests/test_gpu_leaks.py -q -m gpu -x -k "public_api or corpus" > gpurun_out/${tag}_pytest_leaks.log 2>&1; tail -2 gpurun_out/${tag}_pytest_leaks.log
