timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/mgpu_worker.py > gpurun_out/mgpu_worker.log 2>&1; grep -v "^\[W\|^$" gpurun_out/mgpu_worker.log | grep -B2 -A12 "Traceback\|Error\|MGPU" | head -60
timeout 900 python -m pytest tests/test_fft3_gpu.py -x -q 2>&1 | tail -5
for lpc in 8 4; do CHB_Z_LPC=$lpc python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_lpc$lpc.json 2> gpurun_out/bench_c3_lpc$lpc.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c3_lpc$lpc.json')); print('LPC$lpc', d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('fail', e); print(open('gpurun_out/bench_c3_lpc$lpc.err').read()[-2000:])
PY
done
