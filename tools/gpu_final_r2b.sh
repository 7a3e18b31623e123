# round 2, final measurement set on one B200: phased leak variant (tuning), drop-in tests, smoke, both bench arms, ncu of the headline kernel
tag=r2fin
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
for v in ph ph8 p1; do
  f=gpurun_out/${tag}_leak_${v}_30000_7.log
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so timeout 90 python tools/leak_probe.py 30000 7 only > $f 2>&1
  echo "== leak variant $v"; grep -E "leak_calc|leak kernel|leak clocks" $f | cut -c1-330
done
unset POLYCAP_B200_PHASES
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_ph.so timeout 150 python -m pytest tests/test_gpu_leaks.py -q -m gpu -x -k "corpus or small_workspaces" > gpurun_out/${tag}_pytest_phased.log 2>&1; echo "phased variant tests:"; tail -1 gpurun_out/${tag}_pytest_phased.log
timeout 400 python -m pytest tests/test_reference_programs.py tests/test_cython_module.py tests/test_gpu_parity.py tests/test_gpu_multidevice.py -q -m gpu -x > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_conical.json 2> gpurun_out/${tag}_conical.err
timeout 200 python bench.py --workload xos1 --no-subs > gpurun_out/${tag}_xos1.json 2> gpurun_out/${tag}_xos1.err
A="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-subs"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_conical.csv python bench.py $A --ids 2000000 > gpurun_out/${tag}_ncu_launches.log 2>&1
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-subs"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcb_trace2 -s 1 -c 1 -f -o gpurun_out/prof_${tag}_conical python bench.py $B --ids 4000000 > gpurun_out/${tag}_ncu_conical.log 2>&1
for f in ref conical xos1; do python - gpurun_out/${tag}_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    sp=d.get('stat_parity') or {}
    print(sys.argv[1], "%.4e"%d['value'], (d.get('roofline') or {}).get('frac'), "e2e %.4e"%d['e2e']['value'], "cpu", (d.get('cpu_baseline') or {}).get('value'), "nsig", sp.get('n_sigma_max'), sp.get('pass'), (d.get('clocks') or {}).get('reasons'))
    for k,v in (d.get('sub_records') or {}).items(): print("   sub", k, v.get('value'), v.get('roofline_frac'), v.get('error'), (v.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
ls -la gpurun_out/prof_${tag}*
