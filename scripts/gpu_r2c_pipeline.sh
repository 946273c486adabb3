# Round 2: first GPU run of the chunk pipeline (RHS in place, products per chunk, two streams / green contexts). 1 GPU.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/c_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'GB', round(d['config']['device_bytes_per_gpu']/1e9,1), 'e2e', round(d['e2e']['value'],2))
except Exception as e: print('$name fail', e); print(open('gpurun_out/c_$name.err').read()[-1500:])
PY
}
run c3_default 3 A=1
run c3_w20 3 CHB_WORK_GB=20
run c3_lanes2 3 CHB_LANES=2
run c3_lanes2_green80 3 CHB_LANES=2 CHB_GREEN=80 CHB_VERBOSE=1
run c3_lanes2_green112 3 CHB_LANES=2 CHB_GREEN=112 CHB_VERBOSE=1
grep -h "green" gpurun_out/c_c3_lanes2_green*.err | head
run c2_default 2 A=1
