nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -c 2500 gpurun_out/bench_c3_n2.json; tail -5 gpurun_out/bench_c3_n2.err
