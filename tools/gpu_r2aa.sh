# round 2: phased leak kernel, wall march in budgeted stretches per clock tick (LK_W_BUDGET): variants, then parity tests on the default build
tag=r2aa
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
for v in w0 w5 w8 w12; do
  f=gpurun_out/${tag}_leak_${v}_30000_7.log
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so timeout 60 python tools/leak_probe.py 30000 7 only > $f 2>&1
  echo "== leak variant $v"; grep -E "leak kernel|leak clocks" $f | cut -c1-330
done
f=gpurun_out/${tag}_leak_default_20000_400.log
timeout 90 python tools/leak_probe.py 20000 400 only > $f 2>&1
echo "== default build, 20000 400"; grep -E "leak_calc|leak kernel|phases: wave" $f | cut -c1-330
unset POLYCAP_B200_PHASES
timeout 200 python -m pytest tests/test_gpu_leaks.py tests/test_reference_programs.py tests/test_cython_module.py -q -m gpu -x -k "corpus or small_workspaces or public_api or long_grid or fast_wall or leaks_program or cython" > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
