# Round 2, final single-GPU evidence: the bench line as the driver runs it, the reference arm, the ncu launch list of the
# same command and one `ncu --set full` capture per hot kernel (config-3 line lengths on a thin grid; the persistent
# x-pass at the headline's nxd = 1536).  Summaries go to profiles/r2l_*.
set -x
nproc
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/l_bench_c3_n1.json 2> gpurun_out/l_bench_c3_n1.err; tail -c 400 gpurun_out/l_bench_c3_n1.json
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/l_reference_c3.json 2> gpurun_out/l_reference_c3.err; tail -c 900 gpurun_out/l_reference_c3.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l_launches_c3.csv \
    python bench.py --workload 3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/l_ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"xpass4|zbwd4|zfwd4|solve_s|rhs_kernel|mean_mode" -s 13 -c 13 \
    -o gpurun_out/prof_r2l_c3shape python bench.py --workload 511,32,511 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/l_ncu_c3shape.log 2>&1
tail -2 gpurun_out/l_ncu_c3shape.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"xpass4|zbwd4|zfwd4" -s 6 -c 3 \
    -o gpurun_out/prof_r2l_c4shape python bench.py --workload 1023,16,1023 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/l_ncu_c4shape.log 2>&1
tail -2 gpurun_out/l_ncu_c4shape.log
# the y-direction kernels at the per-GPU shape of the headline grid on 8 GPUs (128 x-modes, ny = 1024, nz = 1023)
for v in A=1 CHB_SOLVE_PF=1; do
  env $v timeout 300 python bench.py --workload 127,1024,1023 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/l_yshape_$v.json 2> gpurun_out/l_yshape_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/l_yshape_$v.json')); print('$v', round(d['ms_per_step'],1), {k:round(x['ms_per_step'],1) for k,x in d['kernels'].items()}, d['kernels']['solve'].get('parts_ms_per_step'))
except Exception as e: print('$v fail', e)
PY
done
