# Round 2: the persistent x-pass (inputs of the next line prefetched into shared memory, tables resident) against the
# one-CTA-per-line kernels, with one / two threads per butterfly position, at nxd = 1536 (headline shape) and 768. 1 GPU.
set -x
timeout 600 python -m pytest tests/test_zz_experimental_gpu.py tests/test_pipeline_gpu.py -x -q -k "persistent or chunking or rhs_lives" 2>&1 | tail -5
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/e_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['kernels']['solve'].get('parts_ms_per_step'))
except Exception as e: print('$name fail', e); print(open('gpurun_out/e_$name.err').read()[-1500:])
PY
}
run c4s_default 1023,16,1023 A=1
run c4s_nosplit 1023,16,1023 CHB_XPASS_SPLIT=0
run c4s_persist 1023,16,1023 CHB_XPASS_PERSIST=1
run c4s_persist_nosplit 1023,16,1023 CHB_XPASS_PERSIST=1 CHB_XPASS_SPLIT=0
run c3_default 3 A=1
run c3_persist 3 CHB_XPASS_PERSIST=1
run c3_persist_split 3 CHB_XPASS_PERSIST=1 CHB_XPASS_SPLIT=1
