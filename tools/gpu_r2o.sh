# round 2: leak kernel with its functions out of line (instruction-cache footprint), cheaper jumps: parity, throughput, ncu
tag=r2o
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_leaks.py -q -m gpu -x > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
export POLYCAP_B200_PHASES=1
for a in "1000 400" "30000 400" "100000 7" ; do
  timeout 600 python tools/leak_probe.py $a only > gpurun_out/${tag}_probe_$(echo $a | tr ' ' '_').log 2>&1; grep -E "leak_calc|leak wave" gpurun_out/${tag}_probe_$(echo $a | tr ' ' '_').log | tail -3
done
unset POLYCAP_B200_PHASES
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pcb_leak_kernel -c 1 -f -o gpurun_out/prof_${tag}_leak python tools/leak_probe.py 4000 40 only > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/prof_${tag}*
