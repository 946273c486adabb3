# Round 2, N GPUs (N = 4 or 8): the chunk pipeline with the transposes and the local kernels on disjoint SM partitions
# against the sequential sweep, and the headline leg (config 4).  usage: bash scripts/gpu_r2g_overlap.sh N
N=${1:-4}
set -x
run() { name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 5 --warmup 3 $EXTRA > gpurun_out/g${N}_$name.json 2> gpurun_out/g${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/g${N}_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'nvlink', round(d['nvlink']['step']['frac'],3), 'roof', round(d['step_roofline']['frac_of_max_hbm_nvlink'],3), 'parity', d.get('parity_check',{}).get('worst_rel_err'))
    h=d.get('headline_config4')
    if h: print('  headline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in h.items() if k in ('ran','why','ms_per_step','steps_per_s','ns_per_dof_step','device_bytes_per_gpu')}, h.get('step_roofline'), {k:round(v['ms_per_step'],1) for k,v in h.get('kernels',{}).items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/g${N}_$name.err').read()[-2500:])
PY
}
EXTRA=""
run default CHB_VERBOSE=1
grep -h "green" gpurun_out/g${N}_default.err | head -2
EXTRA="--no-headline --no-parity-check"
run lanes1 CHB_LANES=1
run green0 CHB_GREEN=0
run green96 CHB_GREEN=96
run green64 CHB_GREEN=64
EXTRA="--no-parity-check --workload 4"
run c4_lanes1 CHB_LANES=1
