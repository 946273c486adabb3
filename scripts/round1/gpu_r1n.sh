# round 1, call n: fused y-direction flow + restart I/O on the GPU (tests), fuse variants at config 3, snapshot timing
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_restart_io_gpu.py -x -q 2>&1 | tail -8
run() { # name, env...
  name=$1; shift
  env "$@" timeout 170 python bench.py --workload ${WL:-3} --steps 2 --warmup 1 --no-cpu-baseline ${EXTRA:-} > gpurun_out/n_$name.json 2> gpurun_out/n_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/n_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'e2e', round(d['e2e']['value'],3), d.get('snapshot'))
except Exception as e: print('$name fail', e); print(open('gpurun_out/n_$name.err').read()[-1500:])
PY
}
run fuse1 CHB_FUSE=1
run fuse2 CHB_FUSE=2
run fuse0 CHB_FUSE=0
WL=2 EXTRA=--snapshot run c2_snap CHB_FUSE=1
