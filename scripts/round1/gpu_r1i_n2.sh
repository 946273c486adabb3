set -x
nvidia-smi --query-gpu=name --format=csv | head -3
timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/n2_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/n2_$name.err').read()[-2500:])
PY
}
run p2p A=1
run nccl CHB_P2P=0
