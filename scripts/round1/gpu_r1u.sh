# L2-resident chunks of planes at config 2 (4 planes = 85 MB of pencil-transpose buffers): one bench run
CHB_WORK_GB=0.08 timeout 36 python bench.py --workload 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/u_c2_l2.json 2> gpurun_out/u_c2_l2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/u_c2_l2.json')); print('c2_l2chunks', round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['gpu_launches'])
except Exception as e: print('fail', e); print(open('gpurun_out/u_c2_l2.err').read()[-800:])
PY
