# 8-GPU runs of the final build: the weak-scaling bench line and the 1e9-id XOS run (run under gpurun --gpus 8)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/s8_conical.json 2> gpurun_out/s8_conical.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload xos1 --steps 63 --no-e2e > gpurun_out/s8_xos1_1e9.json 2> gpurun_out/s8_xos1_1e9.err
python bench.py --no-cpu-baseline > gpurun_out/s8_conical_1gpu.json 2> gpurun_out/s8_conical_1gpu.err
for f in gpurun_out/s8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['n_gpus'], "%.4e"%d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
