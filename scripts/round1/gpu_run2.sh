set -x
python -c "
import ctypes as C, numpy as np
from channel_b200 import _lib
lib=_lib.load(); o=np.zeros(2); lib.chb_measure_device_peaks(o.ctypes.data_as(_lib.c_double_p)); print('PEAKS fp64_tflops=%.2f copy_gbs=%.1f'%(o[0],o[1]))
" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
python bench.py --workload 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
