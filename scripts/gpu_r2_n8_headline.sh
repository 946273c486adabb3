# Round 2, 8 GPUs: the headline of BASELINE.json - config 4 (1023x1024x1023) on 8 B200s - never measured in round 1.
# 58 GB of fields per GPU; each rank builds its 12.9 GB slab of the synthetic field on the host first (minutes).
run() { name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 8 --workload 4 --steps 3 --warmup 3 > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/h_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), 'steps/s', round(d['value'],2), 'ns/DoF/step', round(d['ns_per_dof_step'],4), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['nvlink'], 'step_roofline', round(d['step_roofline']['frac'],3))
except Exception as e: print('$name fail', e); print(open('gpurun_out/h_$name.err').read()[-3000:])
PY
}
run config4_default A=1
# after scripts/gpu_r2_variants.sh has shown which experimental kernels win at nxd 1536 / nzd 3072:
run config4_variants CHB_XPASS_SPLIT=1 CHB_Z_TPL=128 CHB_ZF_LPC=2 CHB_ZB_LPC=2
