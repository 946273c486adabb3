set -x
timeout 1500 python -m pytest tests/test_fft3_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --workload ${WL:-3} --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/v_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/v_$name.err').read()[-1500:])
PY
}
run default A=1
run zb4 CHB_ZB_LPC=4
WL=2 run c2_default A=1
