# round 1, call o: bench line (config 3) with snapshot timing, full GPU test suite, rhs occupancy variant, ncu of the y-direction kernels
set -x
python bench.py --steps 4 --warmup 3 --snapshot > gpurun_out/r1e_bench_c3.json 2> gpurun_out/r1e_bench_c3.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r1e_bench_c3.json')); print('c3', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['kernels']['solve'].get('parts_ms_per_step'), d['e2e'], d['roofline'], d.get('snapshot'), d.get('cpu_baseline'))
except Exception as e: print('c3 fail', e); print(open('gpurun_out/r1e_bench_c3.err').read()[-2000:])
PY
timeout 540 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
CHB_RHS_MINB=4 timeout 170 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/o_rhs4.json 2> gpurun_out/o_rhs4.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/o_rhs4.json')); print('rhs4', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
except Exception as e: print('rhs4 fail', e); print(open('gpurun_out/o_rhs4.err').read()[-1500:])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches_c3.csv python bench.py --workload 3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_ncu_launch.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"rhs_kernel|solve_s" -s 5 -c 5 -o gpurun_out/prof_r1e_ydir python bench.py --workload 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1e_ncu_full.log 2>&1
tail -2 gpurun_out/r1e_ncu_full.log
