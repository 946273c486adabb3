# Round 2, 2 GPUs: multi-rank parity of the chunk pipeline (direct NVLink stores, NCCL, one / two lanes, green contexts),
# then the bench line with its parity_check and headline_config4 legs, then the overlap variants.
set -x
nproc; grep -E "MemTotal|MemAvailable" /proc/meminfo
timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "2" 2>&1 | tail -5
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 5 --warmup 3 $EXTRA > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/d_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'nvlink', round(d['nvlink']['step']['frac'],3), 'roof', round(d['step_roofline']['frac_of_max_hbm_nvlink'],3), 'parity', d.get('parity_check',{}).get('worst_rel_err'))
    h=d.get('headline_config4')
    if h: print('  headline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in h.items() if k in ('ran','why','ms_per_step','steps_per_s','ns_per_dof_step','device_bytes_per_gpu')}, h.get('step_roofline'), {k:round(v['ms_per_step'],1) for k,v in h.get('kernels',{}).items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/d_$name.err').read()[-2500:])
PY
}
EXTRA=""
run default CHB_VERBOSE=1
grep -h "green" gpurun_out/d_default.err | head -3
EXTRA="--no-headline --no-parity-check"
run lanes1 CHB_LANES=1
run green0 CHB_GREEN=0
run green96 CHB_GREEN=96
run green64 CHB_GREEN=64
