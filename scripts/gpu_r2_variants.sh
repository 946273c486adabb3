# Round 2, first GPU call: the experimental kernel variants that were proven on the CPU emulator in round 1
#  against the default kernels.  1 GPU; prints ms/step and the per-kernel split of each run.
#   config 3 (511x512x511): the bench workload;  1023,16,1023: the line lengths of the headline grid (nxd 1536, nzd 3072)
set -x
timeout 600 python -m pytest tests/test_zz_experimental_gpu.py -q 2>&1 | tail -5
run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 200 python bench.py --workload $wl --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/v_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('$name fail', e); print(open('gpurun_out/v_$name.err').read()[-1500:])
PY
}
run c3_default 3 A=1
run c3_xsplit 3 CHB_XPASS_SPLIT=1
run c3_zfdirect 3 CHB_ZF_DIRECT=1
run c3_ztpl128 3 CHB_Z_TPL=128 CHB_ZF_LPC=2 CHB_ZB_LPC=2
run c3_ztpl96 3 CHB_Z_TPL=96 CHB_ZF_LPC=2 CHB_ZB_LPC=2
run c3_zb128_only 3 CHB_Z_TPL=128 CHB_ZB_LPC=2
run c3_solvepf 3 CHB_SOLVE_PF=1
run c3_rhschunk 3 CHB_RHS_CHUNKED=1
run c3_rhschunk_lanes2 3 CHB_RHS_CHUNKED=1 CHB_LANES=2
run c4shape_default 1023,16,1023 A=1
run c4shape_xsplit 1023,16,1023 CHB_XPASS_SPLIT=1
run c4shape_ztpl128 1023,16,1023 CHB_Z_TPL=128 CHB_ZF_LPC=2 CHB_ZB_LPC=2
run c4shape_all 1023,16,1023 CHB_XPASS_SPLIT=1 CHB_Z_TPL=128 CHB_ZF_LPC=2 CHB_ZB_LPC=2 CHB_ZF_DIRECT=1
