# Round 2, last call: the GPU suite, smoke() and the bench line on the final tree. 1 GPU.
set -x
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 100 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/p_bench.json')); print('bench', round(d['ms_per_step'],1), round(d['e2e']['value'],2), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['roofline']['frac'], d['roofline']['traffic'], d['config']['device_bytes_per_gpu'])
PY
