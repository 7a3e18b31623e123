# round 2, leak tracer rebuilt (warp per photon, lane-parallel scan, jumped wall march): parity tests, then throughput probes
tag=r2n
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_leaks.py tests/test_reference_programs.py tests/test_cython_module.py -q -m gpu -x > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log
export POLYCAP_B200_PHASES=1
for a in "1000 7" "1000 400" "20000 400" "100000 7" ; do
  timeout 600 python tools/leak_probe.py $a only > gpurun_out/${tag}_probe_$(echo $a | tr ' ' '_').log 2>&1; grep -E "leak_calc|leak wave" gpurun_out/${tag}_probe_$(echo $a | tr ' ' '_').log | tail -4
done
