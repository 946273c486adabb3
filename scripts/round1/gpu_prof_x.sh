ncu --set full --clock-control none --import-source on -k regex:"xpass3" -s 3 -c 1 -o gpurun_out/prof_x3 python bench.py --workload 511,24,511 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_x3.log 2>&1
tail -3 gpurun_out/ncu_x3.log
