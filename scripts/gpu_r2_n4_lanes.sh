# Round 2, 4 GPUs: does the two-lane chunk pipeline hide the local kernels behind the NVLink-bound ones?
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/l_$name.json 2> gpurun_out/l_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/l_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, round(d['nvlink']['step']['frac'],3))
except Exception as e: print('$name fail', e); print(open('gpurun_out/l_$name.err').read()[-2500:])
PY
}
run default A=1
run lanes2 CHB_LANES=2
run lanes2_w4 CHB_LANES=2 CHB_WORK_GB=4
run lanes2_w2 CHB_LANES=2 CHB_WORK_GB=2
run lanes1_w4 CHB_WORK_GB=4
run solvepf CHB_SOLVE_PF=1
run rhschunk_w4 CHB_RHS_CHUNKED=1 CHB_WORK_GB=4
run rhschunk_lanes2_w4 CHB_RHS_CHUNKED=1 CHB_LANES=2 CHB_WORK_GB=4
