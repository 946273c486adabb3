# Round 2, 8 GPUs: chunk size of the pipeline on the headline grid (fill / drain of the two-stream pipeline against launch and
# barrier overhead): 10 GB arena = 61 planes per chunk (17 chunks per sweep, default), 5 GB = 30 planes, 2.5 GB = 15 planes.
N=8
set -x
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --no-parity-check --workload 4 --steps 5 --warmup 2 > gpurun_out/o${N}_$name.json 2> gpurun_out/o${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/o${N}_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'roof', round(d['step_roofline']['frac_of_max_hbm_nvlink'],4), 'GB', round(d['config']['device_bytes_per_gpu']/1e9,1))
except Exception as e: print('$name fail', e); print(open('gpurun_out/o${N}_$name.err').read()[-2500:])
PY
}
run w5 CHB_WORK_GB=5
run w2p5 CHB_WORK_GB=2.5
