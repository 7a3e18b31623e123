# 8 x B200 on one box: the driver's scaling sweep (one process per GPU), the library's own multi-device path from one process, the 1e9-photon XOS run
tag=r2t
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
python -m pytest tests/test_gpu_multidevice.py -q -m gpu -k "real_devices" > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_conical_${N}gpu.json 2> gpurun_out/${tag}_conical_${N}gpu.err
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_conical_1gpu.json 2> gpurun_out/${tag}_conical_1gpu.err
python tools/multi_device_probe.py conical > gpurun_out/${tag}_probe_conical.json 2> gpurun_out/${tag}_probe_conical.err
python tools/multi_device_probe.py xos1 > gpurun_out/${tag}_probe_xos1.json 2> gpurun_out/${tag}_probe_xos1.err
python tools/stream_1e9.py 1e9 > gpurun_out/${tag}_stream_1e9.json 2> gpurun_out/${tag}_stream_1e9.err; cut -c1-600 gpurun_out/${tag}_stream_1e9.json; tail -2 gpurun_out/${tag}_stream_1e9.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --workload xos1 --steps 63 --warmup 3 --no-subs --no-e2e --no-cpu-baseline > gpurun_out/${tag}_xos1_1e9_8gpu.json 2> gpurun_out/${tag}_xos1_1e9_8gpu.err
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/r2t_conical_%dgpu.json"%n).read().strip().splitlines()[-1])
        sp=d.get("single_process_multi_gpu") or {}
        print(n, "value %.4e e2e %.4e ms/step %.2f allreduce %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["tally_allreduce_ms"]),
              "xos1 %.4e"%d["sub_records"]["xos1"]["value"], "ellip %.4e"%d["sub_records"]["ellip"]["value"],
              "1proc:", {k:"%.4e"%sp[k]["value"] for k in ("conical","xos1") if k in sp}, sp.get("error"), d["clocks"]["reasons"])
    except Exception as e: print(n, "ERR", e)
for w in ("conical","xos1"):
    try:
        d=json.loads(open("gpurun_out/r2t_probe_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "probe speedup %.2f"%d["speedup"], [(r["devices"], "%.4e"%r["photons_per_s"], r["seconds_per_call"]) for r in d["runs"]])
    except Exception as e: print(w, "probe ERR", e)
try:
    d=json.loads(open("gpurun_out/r2t_xos1_1e9_8gpu.json").read().strip().splitlines()[-1])
    print("xos1 1e9: ids", d["counters"]["ids"], "value %.4e photons/s, %.2f s device time"%(d["value"], d["ms_per_step"]*d["steps"]/1e3), d["clocks"])
except Exception as e: print("xos 1e9 ERR", e)
PY
