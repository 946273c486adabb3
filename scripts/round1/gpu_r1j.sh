set -x
timeout 1500 python -m pytest tests/test_fft3_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_fft_gpu.py -x -q 2>&1 | tail -8
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --workload ${WL:-3} --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/v_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/v_$name.err').read()[-1500:])
PY
}
run default A=1
run lanes1 CHB_LANES=1
run zf8 CHB_ZF_LPC=8
run twa_rm CHB_TWA=-1
WL=2 run c2_default A=1
WL=2 run c2_lanes1 CHB_LANES=1
