# Round 2: GPU suite (incl. the full-size parity tests against the C oracle), then the L2 prefetch of the z-pass inputs. 1 GPU.
set -x
nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/i_$name.json 2> gpurun_out/i_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/i_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/i_$name.err').read()[-1500:])
PY
}
run c3_default 3 A=1
run c3_pf300 3 CHB_Z_L2PF=300
run c3_pf600 3 CHB_Z_L2PF=600
run c3_pf1200 3 CHB_Z_L2PF=1200
run c3_pf2400 3 CHB_Z_L2PF=2400
run c4s_default 1023,16,1023 A=1
run c4s_pf600 1023,16,1023 CHB_Z_L2PF=600
