timeout 900 python -m pytest tests/test_fft3_gpu.py -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('fail', e); print(open('gpurun_out/bench_c3.err').read()[-2000:])
PY
python bench.py --workload 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c2.json')); print('c2', d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('fail', e); print(open('gpurun_out/bench_c2.err').read()[-2000:])
PY
