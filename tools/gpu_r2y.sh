# round 2: phased leak kernel as the default: occupancy variants, then the leak parity tests and the reference's own programs on the default build
tag=r2y
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
for v in ph3 ph4 ph6w; do
  f=gpurun_out/${tag}_leak_${v}_30000_7.log
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so timeout 90 python tools/leak_probe.py 30000 7 only > $f 2>&1
  echo "== leak variant $v"; grep -E "leak_calc|leak kernel|leak clocks|leak wave" $f | cut -c1-330
done
for a in "30000 7" "20000 400"; do
  f=gpurun_out/${tag}_leak_default_$(echo $a | tr ' ' '_').log
  timeout 120 python tools/leak_probe.py $a only > $f 2>&1
  echo "== default build, $a"; grep -E "leak_calc|leak kernel|phases: wave" $f | cut -c1-330
done
unset POLYCAP_B200_PHASES
timeout 400 python -m pytest tests/test_gpu_leaks.py tests/test_reference_programs.py tests/test_cython_module.py -q -m gpu -x > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
