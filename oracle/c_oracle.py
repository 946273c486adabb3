"""ctypes binding of oracle/channel_oracle_c.c (C99 + OpenMP restatement of the reference).

TEST INFRASTRUCTURE ONLY (see the header of channel_oracle_c.c): used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libchannel_oracle.so")
_lib = None
_dp = C.POINTER(C.c_double)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE])
    lib = C.CDLL(LIB_PATH)
    lib.co_create.restype = C.c_void_p
    lib.co_create.argtypes = [C.c_int] * 3 + [C.c_double] * 6
    lib.co_destroy.argtypes = [C.c_void_p]
    lib.co_set_params.argtypes = [C.c_void_p] + [C.c_double] * 4 + [C.c_int, C.c_int] + [C.c_double] * 6
    lib.co_global_threads.argtypes = [C.c_int]
    lib.co_set_threads.argtypes = [C.c_void_p, C.c_int]
    lib.co_threads.argtypes = [C.c_void_p]
    lib.co_threads.restype = C.c_int
    for n in ("co_V", "co_F", "co_oldrhs"):
        getattr(lib, n).restype = C.c_void_p
        getattr(lib, n).argtypes = [C.c_void_p]
    lib.co_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.co_table.restype = _dp
    lib.co_table.argtypes = [C.c_void_p, C.c_int]
    lib.co_get_scalars.argtypes = [C.c_void_p, _dp]
    lib.co_fft_lines.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.co_set_coriolis.argtypes = [C.c_void_p] + [C.c_double] * 3
    lib.co_set_body_force.argtypes = [C.c_void_p]
    lib.co_set_am.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.co_buildrhs.argtypes = [C.c_void_p, _dp, C.c_int]
    lib.co_linsolve.argtypes = [C.c_void_p, C.c_double]
    lib.co_cfl_prepass.argtypes = [C.c_void_p]
    lib.co_outstats.argtypes = [C.c_void_p, _dp]
    lib.co_step.argtypes = [C.c_void_p, _dp]
    lib.co_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.co_fill_synthetic.argtypes = [C.c_void_p]
    _lib = lib
    return lib


class COracle:
    """Same surface as oracle.channel_oracle.Oracle where the tests need it.  V is exposed in
    the reference layout V(iy,iz,ix,c) = C-order [c][ix][iz+nz][iy+1]."""

    TABLES = dict(d0=0, d1=1, d2=2, d4=3, D0mat=4, y=5, d140=10, d14m1=11, d240=12, d24m1=13, d14n=14,
                  d14np1=15, d24n=16, d24np1=17, v0bc=20, v0m1bc=21, vnbc=22, vnp1bc=23, eta0bc=24,
                  eta0m1bc=25, etanbc=26, etanp1bc=27)

    def __init__(self, p):
        lib = load()
        self.lib, self.p = lib, p
        self.nx, self.ny, self.nz = p.nx, p.ny, p.nz
        self.h = lib.co_create(p.nx, p.ny, p.nz, p.alfa0, p.beta0, 1.0 / p.re, p.a, p.ymin, p.ymax)
        if not self.h:
            raise MemoryError("co_create failed (not enough host memory for this grid)")
        lib.co_set_params(self.h, p.meanpx, p.meanpz, p.meanflowx, p.meanflowz, int(p.CPI), p.CPI_type, p.gamma,
                          p.u0, p.uN, p.deltat, p.cflmax, p.time)
        a, b = C.c_int(), C.c_int()
        lib.co_sizes(self.h, C.byref(a), C.byref(b))
        self.nxd, self.nzd = a.value, b.value
        shape = (3, p.nx + 1, 2 * p.nz + 1, p.ny + 3)
        n = int(np.prod(shape))
        buf = (C.c_double * (2 * n)).from_address(lib.co_V(self.h))
        self.Vf = np.frombuffer(buf, dtype=np.complex128).reshape(shape)

    def table(self, name):
        ny = self.ny
        n = {"D0mat": (ny + 1) * 5, "y": ny + 3}.get(name, (ny + 3) * 5 if name in ("d0", "d1", "d2", "d4") else 5)
        ptr = self.lib.co_table(self.h, self.TABLES[name])
        a = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
        return a.reshape(-1, 5) if n > 5 and name != "y" else a

    # device-layout views [c, iy+1, ix, iz+nz] <-> reference layout
    def set_V(self, V):
        self.Vf[...] = np.transpose(V, (0, 2, 3, 1))

    def get_V(self):
        return np.ascontiguousarray(np.transpose(self.Vf, (0, 3, 1, 2)))

    def scalars(self):
        s = np.zeros(10)
        self.lib.co_get_scalars(self.h, s.ctypes.data_as(_dp))
        return dict(cfl=s[0], fr=s[1:4].copy(), corrpx=s[4], corrpz=s[5], meanpx=s[6], meanpz=s[7], deltat=s[8], time=s[9])

    def set_coriolis(self, omega2, kz_cutoff, y_threshold_bot):
        self.lib.co_set_coriolis(self.h, omega2, kz_cutoff, y_threshold_bot)
        self.lib.co_set_body_force(self.h)

    def set_am(self, which, lambdaz_f, amp):
        """which = "am_f1" | "am_butterfly" (body_forces/am_f1/am_f1.inc, am_butterfly/am_butterfly.inc)"""
        self.lib.co_set_am(self.h, {"am_f1": 2, "am_butterfly": 3}[which], lambdaz_f, amp)
        self.lib.co_set_body_force(self.h)

    def cfl_prepass(self):
        self.lib.co_cfl_prepass(self.h)

    def buildrhs(self, ODE, compute_cfl):
        ode = np.array(ODE, dtype=np.float64)
        self.lib.co_buildrhs(self.h, ode.ctypes.data_as(_dp), int(compute_cfl))

    def linsolve(self, lam):
        self.lib.co_linsolve(self.h, lam)

    def outstats(self):
        line = np.zeros(11)
        self.lib.co_outstats(self.h, line.ctypes.data_as(_dp))
        return line

    def step(self):
        line = np.zeros(11)
        self.lib.co_step(self.h, line.ctypes.data_as(_dp))
        return line

    def close(self):
        if self.h:
            self.Vf = None
            self.lib.co_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fft_lines(x, sign):
    """Unnormalised complex FFT of the rows of x (sign +1 backward / -1 forward)."""
    lib = load()
    y = np.ascontiguousarray(x, dtype=np.complex128).copy()
    lib.co_fft_lines(y.shape[1], y.shape[0], sign, y.ctypes.data)
    return y


_sample_state = {}


def host_cores():
    """cores this process may run on (the affinity mask, not OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def timed_sample(nx, ny, nz, seconds=15.0, threads=None, alfa0=0.5, beta0=1.0, re=12431.0):
    """Time a bounded sample of one RK3 step of the reference algorithm on the host cores and
    scale it to a full step.  Sample = `iters` iterations of buildrhs's y-plane loop (each a
    full-plane `convolutions` plus the RHS assembly of one plane) and the columns of `nix`
    x-wavenumbers of linsolve, of RK substep 2; a full step is 3 substeps of ny+3 convolution
    planes, ny-1 RHS planes and nx+1 x-wavenumbers.  threads=None: every core of the affinity mask."""
    from .channel_oracle import DnsIn
    lib = load()
    threads = int(threads) if threads else host_cores()
    key = (nx, ny, nz, threads)
    if key not in _sample_state:
        _sample_state.clear()
        lib.co_global_threads(threads)
        p = DnsIn(nx=nx, ny=ny, nz=nz, alfa0=alfa0, beta0=beta0, re=re, deltat=1e-3, cflmax=0.0)
        _sample_state[key] = COracle(p)
    o = _sample_state[key]
    lib.co_set_threads(o.h, threads)
    cores = lib.co_threads(o.h)
    out = np.zeros(6)

    def run(iters, nix):
        lib.co_fill_synthetic(o.h)
        lib.co_set_params(o.h, 0.0, 0.0, 0.0, 0.0, 1, 1, 0.161436, 0.0, 0.0, 1e-3, 0.0, 0.0)
        lib.co_sample(o.h, iters, nix, out.ctypes.data_as(_dp))
        return out.copy()

    t0 = time.perf_counter()
    r = run(6, 1)                                   # calibration: 6 conv planes, 2 rhs planes, 1 ix
    per_plane = (r[0] / r[1]) + (r[2] / max(r[3], 1))
    per_ix = r[4] / r[5]
    budget = max(1.0, seconds - (time.perf_counter() - t0))
    iters = int(min(ny + 5, max(8, 0.8 * budget / max(per_plane, 1e-9))))
    nix = int(min(nx + 1, max(1, 0.2 * budget / max(per_ix, 1e-9))))
    t1 = time.perf_counter()
    r = run(iters, nix)
    wall = time.perf_counter() - t1
    t_sub = r[0] / r[1] * (ny + 3) + r[2] / max(r[3], 1) * (ny - 1) + r[4] / r[5] * (nx + 1)
    t_step = 3.0 * t_sub
    frac = (r[0] + r[2] + r[4]) / t_step            # share of a full step's CPU work that was actually executed
    return {"value": 1.0 / t_step, "unit": "steps/s", "cores": int(cores), "kind": "port",
            "sample": f"{int(r[1])} of {ny + 3} convolution planes + {int(r[3])} of {ny - 1} RHS planes + "
                      f"{int(r[5])} of {nx + 1} x-wavenumbers of linsolve (RK substep 2), scaled to 3 substeps; "
                      f"C99+OpenMP restatement of the reference algorithm with its own Stockham FFT (no FFTW/MPI/"
                      f"gfortran in this image), {time.perf_counter() - t0:.1f} s of CPU work on {int(cores)} threads",
            "extrapolated": True, "sampled_fraction_of_step": frac, "sample_wall_s": wall,
            "seconds_per_step_est": t_step,
            "split_s_per_substep": {"convolutions": r[0] / r[1] * (ny + 3), "rhs": r[2] / max(r[3], 1) * (ny - 1),
                                    "linsolve": r[4] / r[5] * (nx + 1)}}


def timed_full_step(nx, ny, nz, threads=None, alfa0=0.5, beta0=1.0, re=12431.0):
    """One COMPLETE, unsampled RK3 step of the reference algorithm (co_step: 3 x buildrhs + linsolve, outstats) on the
    host cores, next to what timed_sample extrapolates for the same grid: the check of the extrapolation."""
    from .channel_oracle import DnsIn
    lib = load()
    threads = int(threads) if threads else host_cores()
    lib.co_global_threads(threads)
    p = DnsIn(nx=nx, ny=ny, nz=nz, alfa0=alfa0, beta0=beta0, re=re, deltat=1e-3, cflmax=0.0)
    o = COracle(p)
    lib.co_fill_synthetic(o.h)
    lib.co_set_params(o.h, 0.0, 0.0, 0.0, 0.0, 1, 1, 0.161436, 0.0, 0.0, 1e-3, 0.0, 0.0)
    line = np.zeros(11)
    lib.co_step(o.h, line.ctypes.data_as(_dp))      # warm-up: first touch of the work arrays
    t0 = time.perf_counter()
    lib.co_step(o.h, line.ctypes.data_as(_dp))
    t_full = time.perf_counter() - t0
    cores = lib.co_threads(o.h)
    o.close()
    est = timed_sample(nx, ny, nz, seconds=min(8.0, max(2.0, t_full)), threads=threads, alfa0=alfa0, beta0=beta0, re=re)
    return {"grid": [nx, ny, nz], "cores": int(cores), "full_step_s": t_full, "finite": bool(np.isfinite(line).all()),
            "sampled_estimate_s": est["seconds_per_step_est"], "sampled_fraction_of_step": est["sampled_fraction_of_step"],
            "estimate_over_full": est["seconds_per_step_est"] / t_full}
