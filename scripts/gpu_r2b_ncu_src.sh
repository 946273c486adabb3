# Round 2: source-level ncu capture (--import-source on, -lineinfo build) of one launch of xpass4 / zbwd4 / zfwd4 /
# solve_s2 / rhs at the line lengths of config 3 (nxd 768, nzd 1536) and of xpass4 at the headline's nxd 1536,
# on thin grids (few planes) so that the ~40 replays per kernel stay short.  1 GPU.
set -x
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"xpass4|zbwd4|zfwd4|solve_s2|rhs_kernel" -s 10 -c 7 \
    -o gpurun_out/prof_r2b_c3shape python bench.py --workload 511,32,511 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_c3shape.log 2>&1
tail -2 gpurun_out/r2b_ncu_c3shape.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"xpass4|zbwd4|zfwd4" -s 6 -c 3 \
    -o gpurun_out/prof_r2b_c4shape python bench.py --workload 1023,16,1023 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_c4shape.log 2>&1
tail -2 gpurun_out/r2b_ncu_c4shape.log
ls -la gpurun_out/*.ncu-rep
