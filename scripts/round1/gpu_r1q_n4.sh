# round 1, call q (4 GPUs, 1 minute): bench at N=4 (first run of the 4-rank direct-NVLink path)
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/q_n4.json 2> gpurun_out/q_n4.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/q_n4.json') if l.startswith('{')][-1]); print('n4', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['nvlink'], d['e2e']['value'], d['finite'])
except Exception as e: print('n4 fail', e); print(open('gpurun_out/q_n4.err').read()[-3000:])
PY
