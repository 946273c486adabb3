set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 4 --warmup 3 > gpurun_out/r1c_bench_c3.json 2> gpurun_out/r1c_bench_c3.err; python - <<PY
import json
d=json.load(open('gpurun_out/r1c_bench_c3.json')); print('c3', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['e2e'], d['roofline'])
PY
CHB_WORK_GB=3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/v_work3.json 2>/dev/null; python - <<PY
import json
d=json.load(open('gpurun_out/v_work3.json')); print('work3', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
PY
python bench.py --workload 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_c2.json 2> gpurun_out/r1c_bench_c2.err; python - <<PY
import json
d=json.load(open('gpurun_out/r1c_bench_c2.json')); print('c2', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_c3.csv python bench.py --workload 3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"xpass|zfwd|zbwd|rhs_kernel|solve_s" -s 12 -c 12 -o gpurun_out/prof_r1c python bench.py --workload 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1c_ncu_full.log 2>&1
tail -2 gpurun_out/r1c_ncu_full.log
