# Round 2: the other in-place y-direction kernels (S1, S4, rhs) with L2-only loads / stores, as S2 (r2m). 1 GPU.
set -x
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n_$name.json 2> gpurun_out/n_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/n_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, {k:round(v,2) for k,v in d['kernels']['solve'].get('parts_ms_per_step').items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/n_$name.err').read()[-1500:])
PY
}
run base 3 A=1
run ycg 3 CHB_Y_CG=1
run rhscg 3 CHB_RHS_CG=1
run both 3 CHB_Y_CG=1 CHB_RHS_CG=1
CHB_Y_CG=1 CHB_RHS_CG=1 timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "one_substep" 2>&1 | tail -2
