# Round 2: the two ncu passes of B200_PROFILING.md on the bench command (1 GPU), after the variants have been chosen.
# Set the winning switches in the environment of the gpurun call (e.g. CHB_XPASS_SPLIT=1 ...); copy the summaries that
# matter into profiles/ (r2a_*), the .ncu-rep stays in gpurun_out/.
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
tail -c 600 gpurun_out/r2a_bench_c3.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_c3.csv \
    python bench.py --workload 3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"xpass|zfwd|zbwd|rhs_kernel|solve_s|mean_mode" -s 14 -c 14 \
    -o gpurun_out/prof_r2a python bench.py --workload 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
tail -2 gpurun_out/r2a_ncu_full.log
# read here with:  ncu -i gpurun_out/prof_r2a.ncu-rep --page raw --csv > /tmp/raw.csv   (see profiles/README.md for the columns kept)
