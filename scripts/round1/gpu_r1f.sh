set -x
timeout 1200 python -m pytest tests/test_fft3_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -8
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
run x4 CHB_X_VAR=4
run pf64 CHB_PF_DIST=64
run tw3 CHB_TW=3
WL=2 run c2_default A=1
ncu --set full --clock-control none --import-source on -k regex:"xpass6" -s 1 -c 1 -o gpurun_out/prof_r1f python bench.py --workload 511,24,511 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1f_ncu_full.log 2>&1
tail -2 gpurun_out/r1f_ncu_full.log
