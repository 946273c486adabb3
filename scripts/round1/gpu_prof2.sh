ncu --set full --clock-control none --import-source on -k regex:"xpass3|zfwd3|zbwd3" -s 9 -c 3 -o gpurun_out/prof_fft3 python bench.py --workload 511,24,511 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fft3.log 2>&1
tail -3 gpurun_out/ncu_fft3.log
