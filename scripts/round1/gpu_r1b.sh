set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload 3 --steps 4 --warmup 3 > gpurun_out/r1b_bench_c3.json 2> gpurun_out/r1b_bench_c3.err; tail -c 4000 gpurun_out/r1b_bench_c3.json; tail -5 gpurun_out/r1b_bench_c3.err
python bench.py --workload 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_bench_c2.json 2> gpurun_out/r1b_bench_c2.err; tail -c 3000 gpurun_out/r1b_bench_c2.json; tail -5 gpurun_out/r1b_bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_c3.csv python bench.py --workload 3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"xpass|zfwd|zbwd|rhs_kernel|solve_s" -s 40 -c 10 -o gpurun_out/prof_r1b python bench.py --workload 511,24,511 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
tail -3 gpurun_out/r1b_ncu_full.log
ls -la gpurun_out
