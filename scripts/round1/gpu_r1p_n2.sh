# round 1, call p (2 GPUs): multi-GPU parity incl. the shared restart file, bench at N=2 with the nvlink object (direct and NCCL modes)
set -x
timeout 300 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -5
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/p_$name.json 2> gpurun_out/p_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/p_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['nvlink'], d['e2e']['value'])
except Exception as e: print('$name fail', e); print(open('gpurun_out/p_$name.err').read()[-2500:])
PY
}
run n2_direct A=1
run n2_nccl CHB_P2P=0
