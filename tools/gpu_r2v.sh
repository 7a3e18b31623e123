# round 2: leak tracer with the multi-segment jump and the pre-filtered scan for the next bore: parity, phase clocks, throughput
tag=r2v
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_leaks.py -q -m gpu -x > gpurun_out/${tag}_pytest_leaks.log 2>&1; tail -3 gpurun_out/${tag}_pytest_leaks.log
export POLYCAP_B200_PHASES=1
for a in "30000 7" "2000 400"; do
  f=gpurun_out/${tag}_p1_$(echo $a | tr ' ' '_').log
  POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_p1.so timeout 300 python tools/leak_probe.py $a only > $f 2>&1
  echo "== profile build, $a"; grep -E "leak_calc|leak kernel|leak clocks" $f | cut -c1-400
done
for a in "30000 7" "20000 400" "100000 7"; do
  f=gpurun_out/${tag}_leak_$(echo $a | tr ' ' '_').log
  timeout 300 python tools/leak_probe.py $a only > $f 2>&1
  echo "== leak $a"; grep -E "leak_calc|leak kernel|phases: wave" $f | cut -c1-330
done
