# Round 2, 8 GPUs: multi-rank parity (all modes), the bench line with its parity_check and headline_config4 legs, and the
# alternatives (two lanes on config 3, one lane on config 4).
N=8
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_worker.py > gpurun_out/j${N}_mgpu.out 2> gpurun_out/j${N}_mgpu.err; echo rc=$?; tail -2 gpurun_out/j${N}_mgpu.out; grep -iE "error|assert" gpurun_out/j${N}_mgpu.err | tail -5
run() { name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > gpurun_out/j${N}_$name.json 2> gpurun_out/j${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/j${N}_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'nvlink', round(d['nvlink']['step']['frac'],3), 'roof', round(d['step_roofline']['frac_of_max_hbm_nvlink'],3), 'parity', d.get('parity_check',{}).get('worst_rel_err'), 'e2e', round(d['e2e']['value'],2))
    h=d.get('headline_config4')
    if h: print('  headline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in h.items() if k in ('ran','why','ms_per_step','steps_per_s','ns_per_dof_step','device_bytes_per_gpu')}, h.get('step_roofline',{}).get('frac_of_max_hbm_nvlink'), {k:round(v['ms_per_step'],1) for k,v in h.get('kernels',{}).items()}, h.get('nvlink',{}).get('zTOx'), h.get('nvlink',{}).get('xTOz'))
except Exception as e: print('$name fail', e); print(open('gpurun_out/j${N}_$name.err').read()[-2500:])
PY
}
EXTRA=""
run default CHB_VERBOSE=1
grep -h "green" gpurun_out/j${N}_default.err | head -1
EXTRA="--no-headline --no-parity-check"
run c3_lanes2 CHB_LANES=2
EXTRA="--no-parity-check --workload 4 --steps 5"
run c4_lanes1 CHB_LANES=1
