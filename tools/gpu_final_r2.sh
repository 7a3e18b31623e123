# round 2 measurement set on one B200: both bench arms, every workload with e2e / cpu baseline / 3-sigma check, C5 sweep, ncu launch list and full captures
tag=${1:-r2fin}
mkdir -p gpurun_out
cp polycap_b200/libpolycap_b200.so gpurun_out/${tag}_lib.so
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
for w in xos1 ellip dubfoc parab monocap; do
  python bench.py --workload $w --no-subs > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
done
python bench.py --workload dubfoc --sweep --no-subs --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/${tag}_c5_sweep.json 2> gpurun_out/${tag}_c5_sweep.err
A="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
python bench.py $A --ids 2000000 > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_conical.csv python bench.py $A --ids 2000000 > gpurun_out/${tag}_ncu_launches.log 2>&1
python bench.py $A --workload xos1 > gpurun_out/${tag}_plain_xos1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_xos1.csv python bench.py $A --workload xos1 > gpurun_out/${tag}_ncu_launches_xos1.log 2>&1
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-subs"
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_xos python bench.py $B --workload xos1 > gpurun_out/${tag}_ncu_xos1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_conical python bench.py $B --ids 4000000 > gpurun_out/${tag}_ncu_conical.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_dubfoc python bench.py $B --workload dubfoc > gpurun_out/${tag}_ncu_dubfoc.log 2>&1
for f in ref conical xos1 ellip dubfoc parab monocap; do python - gpurun_out/${tag}_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    sp=d.get('stat_parity') or {}
    print(sys.argv[1], "%.4e"%d['value'], (d.get('roofline') or {}).get('frac'), "e2e %.4e"%d['e2e']['value'], "cpu", (d.get('cpu_baseline') or {}).get('value'), "nsig", sp.get('n_sigma_max'), sp.get('pass'), (d.get('clocks') or {}).get('reasons'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2fin_c5_sweep.json").read().strip().splitlines()[-1])
    for r in d["c5_sweep"]: print({k:(round(v,5) if isinstance(v,float) else v) for k,v in r.items() if k!="stat_parity"}, (r.get("stat_parity") or {}).get("n_sigma_max"), (r.get("stat_parity") or {}).get("pass"))
except Exception as e: print("sweep ERR", e)
PY
ls -la gpurun_out/prof_${tag}*
