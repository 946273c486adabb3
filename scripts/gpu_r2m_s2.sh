# Round 2: cache operators of the in-place back-substitution sweep S2 (it lost 3 ms/step against round 1's out-of-place one
# at identical SASS and DRAM bytes; ncu: L2 / L1 hit rates 1.5 / 22 % -> 44 / 45 %, long-scoreboard 2.4 -> 3.7). 1 GPU.
set -x
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m_$name.json 2> gpurun_out/m_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, {k:round(v,2) for k,v in d['kernels']['solve'].get('parts_ms_per_step').items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/m_$name.err').read()[-1500:])
PY
}
run memv0 3 CHB_S2_MEMV=0
run memv1 3 CHB_S2_MEMV=1
run memv2 3 CHB_S2_MEMV=2
run memv3 3 CHB_S2_MEMV=3
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "one_substep" 2>&1 | tail -2
