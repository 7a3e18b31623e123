# round 2: last sanity check of the final default build (leak corpus parity, one leak probe)
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_leaks.py -q -m gpu -x -k "corpus or small_workspaces" > gpurun_out/r2ab_pytest.log 2>&1; tail -1 gpurun_out/r2ab_pytest.log
POLYCAP_B200_PHASES=1 timeout 40 python tools/leak_probe.py 30000 7 only 2>&1 | grep -E "leak kernel|leak_calc" | cut -c1-200
