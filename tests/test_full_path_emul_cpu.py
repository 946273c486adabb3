"""The whole C ABI path on the CPU: tests/host_emul/build_full_emul.sh compiles the library's OWN sources (chb_api.cu,
every launcher and kernel, restart_io.cu, host_tables.cpp) with g++, runs the kernels on the CTA emulator
(cta_emul.hpp) and replaces the CUDA runtime by host memory (fake_cudart.cpp).  A few `-m gpu` tests are then run,
unchanged, against that build in a subprocess (tests/host_emul/run_gpu_tests_emulated.py).

This is test infrastructure: it checks the host logic and the kernel logic of the product sources end to end
(handle creation and destruction, launch order and arguments, work-buffer layouts between the passes, restart
files) before a GPU minute is spent.  The product never loads the emulated build; the real `-m gpu` run on a
B200 is what counts for parity and the only place where performance exists."""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMUL = os.path.join(HERE, "host_emul")
LIB = os.path.join(EMUL, "_build", "libchannel_b200_emul.so")


@pytest.fixture(scope="module")
def emulated_library():
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("needs g++ and the CUDA headers")
    csrc = os.path.join(ROOT, "channel_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps += [os.path.join(EMUL, f) for f in ("cta_emul.hpp", "fake_cudart.cpp", "emul_prelude.hpp", "build_full_emul.sh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["sh", os.path.join(EMUL, "build_full_emul.sh")], stdout=subprocess.DEVNULL)
    return LIB


def run_emulated(args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(EMUL, "run_gpu_tests_emulated.py")] + args, cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout[-3000:] + r.stderr[-2000:]


def test_parity_and_layout_tests_on_the_emulated_library(emulated_library):
    """One RK3 step against the oracle on two of the minimal grids, the rejected sizes and the Fortran-layout round
    trip, through chb_create ... chb_destroy of the emulated build."""
    rc, out = run_emulated(["tests/test_parity_gpu.py", "-x", "-q", "-k",
                            "minimal_grids and (2-9-1 or 5-8-4) or rejected_sizes or fortran_layout"])
    assert rc == 0, out
    assert " passed" in out and "failed" not in out


def test_restart_files_on_the_emulated_library(emulated_library):
    """Snapshot files (blocking and asynchronous, ragged chunks) and the restart read with its header check: the
    writer threads, the chunk arithmetic and the byte layout, on host memory."""
    rc, out = run_emulated(["tests/test_restart_io_gpu.py", "-x", "-q", "-k", "byte_identical and 7-16-5 or roundtrip_and_header"])
    assert rc == 0, out
    assert " passed" in out and "failed" not in out


def test_experimental_variants_on_the_emulated_library(emulated_library):
    """Two of the not-yet-measured variants through the C ABI of the emulated build: the split x-pass at nxd = 768 against the
    default kernel and the oracle, the host-side body-force path (chb_upload_F) and the convection-velocity
    diagnostic (convvel.cu).  The others
    (tests/test_zz_experimental_gpu.py, 13 tests) take a quarter of an hour on the emulator and are run by hand:
    python tests/host_emul/run_gpu_tests_emulated.py tests/test_zz_experimental_gpu.py"""
    rc, out = run_emulated(["tests/test_zz_experimental_gpu.py", "-x", "-q", "-k",
                            "host_side_body_force or (xpass_split and 511-8-4) or convection_velocity"])
    assert rc == 0, out
    assert "3 passed" in out and "failed" not in out


def test_cpp_driver_on_the_emulated_library(emulated_library, tmp_path):
    """channel_main.cpp (the C++ restatement of PROGRAM channel) linked against the emulated build: a run from dns.in
    alone writes Runtimedata, the dt_field snapshot and the final Dati.cart.out; a second run restarts from them."""
    import numpy as np
    exe = os.path.join(EMUL, "_build", "channel_b200_run_emul")
    src = os.path.join(ROOT, "channel_b200", "csrc", "channel_main.cpp")
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe) or os.path.getmtime(LIB) > os.path.getmtime(exe):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-o", exe, src, "-L" + os.path.dirname(LIB), "-lchannel_b200_emul",
                               "-Wl,-rpath," + os.path.dirname(LIB), "-pthread"])
    (tmp_path / "dns.in").write_text("""7 16 5   ! nx ny nz
0.5 1.0
1500
1.5 0.0 2.0
.FALSE. 1 0.161436
0.002 0.0
0.0 0.0
0.0 0.0
0.05 0.0 0.0
0.12 -1 1000 .TRUE.
3
1
""")
    r = subprocess.run([exe, "--dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rtd = np.array([[float(x) for x in l.split()] for l in open(tmp_path / "Runtimedata") if l.strip()])
    assert rtd.shape == (4, 11) and np.isfinite(rtd).all()
    assert np.allclose(rtd[:, 0], [0.0, 0.05, 0.10, 0.15]) and abs(rtd[-1, 5] - 2.0) < 1e-3      # time, flow rate of the parabola
    size = 68 + 16 * 3 * 8 * 11 * 19
    assert os.path.getsize(tmp_path / "Dati.cart.out") == size and os.path.getsize(tmp_path / "Dati.cart.1.out") == size
    r = subprocess.run([exe, "--dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "starting from time" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    rtd2 = np.array([[float(x) for x in l.split()] for l in open(tmp_path / "Runtimedata") if l.strip()])
    assert rtd2.shape == (3 + 1 + 3, 11) and np.array_equal(rtd2[:3], rtd[:3]) and np.all(np.diff(rtd2[:, 0]) > 0)


BENCH_ON_EMULATION = r'''
import sys, json, io, contextlib
sys.path.insert(0, sys.argv[1])
import channel_b200._lib as L
L.LIB_PATH = sys.argv[2]
import torch
torch.cuda.is_available = lambda: True            # the emulated build has no device: stand-ins for the three torch.cuda calls
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
import bench
sys.argv = ["bench.py", "--workload", "7,16,5", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--snapshot", "--snapshot-dir", sys.argv[3]]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
print([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
'''


def test_bench_b200_arm_on_the_emulated_library(emulated_library, tmp_path):
    """bench.py's B200 arm from argument parsing to the JSON line (upload, CFL pre-pass, warm-up, the timed steps with
    the per-kernel timers, the end-to-end leg, the snapshot timing, the report) against the emulated build; the
    numbers mean nothing here, the code path and the keys of the driver contract do."""
    import json
    r = subprocess.run([sys.executable, "-c", BENCH_ON_EMULATION, ROOT, emulated_library, str(tmp_path)], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["metric"] == "rk3_timesteps_per_s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["finite"] is True
    assert d["gpu_launches"] > 0 and set(d["kernels"]) >= {"zfwd", "xpass", "zbwd", "rhs", "solve"}
    assert {"roofline", "step_roofline", "e2e", "clocks", "config", "snapshot", "runtimedata_last"} <= set(d)
    assert d["snapshot"]["bytes"] == 3 * 8 * 11 * 19 * 16 and abs(d["runtimedata_last"][5] - 2.0) < 1e-2


def test_smoke_entry_point_on_the_emulated_library(emulated_library):
    """__graft_entry__.smoke() (one RK3 step of config 1 against the oracle) on the emulated build."""
    code = ("import sys; sys.path.insert(0, sys.argv[1]); import channel_b200._lib as L; L.LIB_PATH = sys.argv[2]; "
            "import __graft_entry__ as g; g.smoke()")
    r = subprocess.run([sys.executable, "-c", code, ROOT, emulated_library], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
