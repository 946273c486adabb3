# Round 2, 2 GPUs: the multi-rank parity worker with its full output (a case failed in gpu_r2d_n2.sh), then the bench line
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_worker.py > gpurun_out/f_mgpu.out 2> gpurun_out/f_mgpu.err
echo rc=$?
tail -5 gpurun_out/f_mgpu.out; grep -v "^$" gpurun_out/f_mgpu.err | tail -40
