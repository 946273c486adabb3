set -x
timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
  python - <<PY
import json
try:
    t=open('gpurun_out/n2_$name.json').read().split('\n')
    d=json.loads([l for l in t if l.startswith('{')][0]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/n2_$name.err').read()[-2500:])
PY
}
run default A=1
run lanes1 CHB_LANES=1
run zf8rm CHB_ZF_LPC=8 CHB_TWA=-1
run zf4tw3 CHB_TWA=3
run zf4rm CHB_TWA=-1
run zf8rm_l1 CHB_ZF_LPC=8 CHB_TWA=-1 CHB_LANES=1
