tag=r2k
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
( time python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_leaks.py::test_single_step_entry_points_with_leaks_match_the_reference tests/test_reference_programs.py -q -m gpu ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/${tag}_bench2.json 2> gpurun_out/${tag}_bench2.err
tail -5 gpurun_out/${tag}_bench2.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2k_bench2.json").read().strip().splitlines()[0])
    print("N=2 value %.4e frac %.4f e2e %.4e allreduce_ms %.3f" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["tally_allreduce_ms"]))
    for k,v in d.get("sub_records",{}).items(): print(k, {a:v.get(a) for a in ("value","roofline_frac","error")})
    sp=d.get("single_process_multi_gpu",{})
    print("single process:", sp.get("what","")[:60], sp.get("error"))
    for k in ("conical","xos1"):
        if k in sp: print(k, {a:sp[k].get(a) for a in ("value","n_devices","seconds_per_call","record_download_GBps")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2k_bench2.err").read()[-3000:])
PY
