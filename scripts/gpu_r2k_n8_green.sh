# Round 2, 8 GPUs: SM split of the chunk pipeline on the headline grid (the z-pass partition was the longer one at 80 + 68).
N=8
set -x
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N $EXTRA > gpurun_out/k${N}_$name.json 2> gpurun_out/k${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/k${N}_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'roof', round(d['step_roofline']['frac_of_max_hbm_nvlink'],3), 'e2e', round(d['e2e']['value'],2))
except Exception as e: print('$name fail', e); print(open('gpurun_out/k${N}_$name.err').read()[-2500:])
PY
}
EXTRA="--no-parity-check --workload 4 --steps 5 --warmup 2"
run c4_green72 CHB_GREEN=72
run c4_green64 CHB_GREEN=64
EXTRA="--no-parity-check --no-headline --steps 10 --warmup 3"
run c3_lanes2_green72 CHB_LANES=2 CHB_GREEN=72
