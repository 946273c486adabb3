python bench.py --steps 4 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
python bench.py --workload 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
