"""The y-direction CUDA kernels (rhs_kernel.cu, solve_kernels.cu) run on the CPU.

tests/host_emul/ydir_emul.cpp compiles the kernel SOURCE with g++ (they are one thread per
wavenumber column, no shared memory / shuffles / atomics) and runs it thread by thread.  This checks,
without a GPU, the flow rhs -> S1 -> S2 -> mean mode -> S3 -> S4 against the numpy oracle's buildrhs + linsolve
(dnsdata.f90:611-673, linsolve_blocking.inc:3-107) to the north-star single-step tolerance, 1e-12
relative.  The GPU parity tests run the same comparisons on the device (tests/test_parity_gpu.py).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from channel_b200 import RK1_rai, RK2_rai, RK3_rai
from channel_b200.fields import perturbed_laminar
from oracle.channel_oracle import DnsIn as ODnsIn, Oracle, am_butterfly_force, am_f1_force, coriolis_force

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "ydir_emul.cpp")
BUILD = os.path.join(HERE, "host_emul", "_build")
CUDA_INC = "/usr/local/cuda/include"


def relerr(a, b):
    n = float(np.abs(b).max())
    d = float(np.abs(a - b).max())
    return d / n if n > 0 else d


@pytest.fixture(scope="module")
def emul():
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libydir_emul.so")
    deps = [SRC, os.path.join(HERE, "host_emul", "cta_emul.hpp")] + [os.path.join(HERE, "..", "channel_b200", "csrc", f)
                    for f in ("rhs_kernel.cu", "solve_kernels.cu", "bodyforce_kernels.cu", "solve_device.cuh", "chb_internal.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-pthread",
                               # the kernels' host names also exist in libchannel_b200.so (loaded RTLD_GLOBAL by
                               # other tests): bind this library's references to its own definitions
                               "-fvisibility=hidden", "-Wl,-Bsymbolic", "-I" + CUDA_INC, "-o", so, SRC])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.chb_emul_ydir_substep.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [dp] * 13 + [C.c_double] * 4 + [C.c_int]
    lib.chb_emul_ydir_substep.restype = C.c_int
    lib.chb_emul_body_force.argtypes = [C.c_int] * 3 + [dp] * 6 + [C.c_int, C.c_int, dp]
    lib.chb_emul_body_force.restype = C.c_int
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def run_substep(lib, o, V, P, F, oldrhs, ODE, fused):
    """one substep of the kernels on copies of the oracle's current state; returns V, oldrhs, rhs, scalars"""
    ny = o.ny
    rows = lambda a: np.ascontiguousarray(a[2:ny + 1], dtype=np.float64)
    bc = np.ascontiguousarray(np.concatenate([np.asarray(getattr(o, n), dtype=np.float64) for n in (
        "d140", "d14m1", "d240", "d24m1", "d14n", "d14np1", "d24n", "d24np1",
        "v0bc", "v0m1bc", "vnbc", "vnp1bc", "eta0bc", "eta0m1bc", "etanbc", "etanp1bc")]))
    y = np.ascontiguousarray(o.y, dtype=np.float64)
    d0, d1, d2, d4 = rows(o.d0), rows(o.d1), rows(o.d2), rows(o.d4)
    D0 = np.ascontiguousarray(o.D0mat, dtype=np.float64)
    V = np.ascontiguousarray(V.copy()); P = np.ascontiguousarray(P); oldrhs = np.ascontiguousarray(oldrhs.copy())
    F = np.ascontiguousarray(F) if F is not None else None
    rhs = np.zeros_like(oldrhs)
    sc = np.array([o.meanpx, o.meanpz, o.meanflowx, o.meanflowz, o.gamma, o.p.u0, o.p.uN,
                   float(o.CPI), float(o.CPI_type), 0.0] + [0.0] * 20)
    rc = lib.chb_emul_ydir_substep(o.nx, o.ny, o.nz, o.alfa0, o.beta0, o.ni, _dp(y), _dp(d0), _dp(d1), _dp(d2), _dp(d4),
                                   _dp(bc), _dp(D0), _dp(V.view(np.float64)), _dp(P.view(np.float64)),
                                   _dp(F.view(np.float64)) if F is not None else None,
                                   _dp(oldrhs.view(np.float64)), _dp(rhs.view(np.float64)), _dp(sc),
                                   ODE[0], ODE[1], ODE[2], o.deltat, fused)
    assert rc == 0
    return V, oldrhs, rhs, sc


CASES = [
    dict(nx=6, ny=16, nz=5),                                     # smallest: one-and-a-bit checkpoint blocks
    dict(nx=9, ny=41, nz=7, CPI=False, meanflowx=2.0),           # ny-1 a multiple of 8, constant-flow-rate branch
    dict(nx=5, ny=30, nz=4, couette=True),                       # Couette walls + coriolis body force
    dict(nx=4, ny=24, nz=6, alfa0=0.8, beta0=1.7, a=2.0, ymin=-1.0, ymax=1.0, CPI=True, CPI_type=0, gamma=0.3),   # another box, grid stretching and CPI law
]


@pytest.mark.parametrize("case", CASES)
def test_ydir_kernels_match_oracle(emul, case):
    case = dict(case)
    couette = case.pop("couette", False)
    kw = dict(re=1500.0, deltat=2e-3, cflmax=0.0)
    kw.update(case)
    if couette:
        kw.update(CPI=False, u0=-1.0, uN=1.0)
    p = ODnsIn(**kw)
    o = Oracle(p)
    o.V[:] = perturbed_laminar(p.nx, p.ny, p.nz, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, eps=5e-2, couette=couette)
    if couette:
        o.set_body_force(coriolis_force(0.02, 9999999.0, 1.0))
    o.cfl_prepass(); o.outstats()
    rng = np.random.default_rng(5)
    o.oldrhs[:, 2:p.ny + 1] = 1e-3 * (rng.standard_normal(o.oldrhs[:, 2:p.ny + 1].shape) + 0j)   # exercise ODE(3)
    for ODE in (RK1_rai, RK2_rai, RK3_rai):
        if getattr(o, "_body_force", None) is not None:
            o._body_force(o)
        V0 = o.V.copy(); old0 = o.oldrhs.copy()
        P = o.convolutions(o.V, False)[..., o.izd]
        mp0 = (o.meanpx, o.meanpz)
        rhs_ref = o.buildrhs(ODE, False)           # also extends F to the ghost nodes in place
        F = o.F.copy() if o.F is not None else None
        o.linsolve(ODE[0] / o.deltat)
        mp1 = (o.meanpx, o.meanpz)
        o.meanpx, o.meanpz = mp0                   # the kernels start from the pre-substep scalars
        sl = slice(2, p.ny + 1)
        _, _, rhsk, _ = run_substep(emul, o, V0, P, F, old0, ODE, -1)      # rhs_kernel alone
        for c in range(2):
            assert relerr(rhsk[c, sl], rhs_ref[c, sl]) < 1e-12, ("rhs", c)
        _, old_c, rhs_c, _ = run_substep(emul, o, V0, P, F, old0, ODE, -2)     # chunked march: bit-identical
        _, old_1, _, _ = run_substep(emul, o, V0, P, F, old0, ODE, -1)
        assert np.array_equal(rhs_c, rhsk) and np.array_equal(old_c, old_1)
        V_default = None
        for fused in (0, 2, 3):          # 0: default kernels, 2: chunked rhs march, 3: prefetching S1 / S3 / S4
            Vk, oldk, rhsk, sc = run_substep(emul, o, V0, P, F, old0, ODE, fused)
            if fused == 0:
                V_default = Vk
            else:
                assert np.array_equal(Vk, V_default), fused      # same operations in the same order
            for c in range(3):
                assert relerr(Vk[c], o.V[c]) < 1e-12, (fused, "field", c, relerr(Vk[c], o.V[c]))
            for c in range(2):
                assert relerr(oldk[c, sl], o.oldrhs[c, sl]) < 1e-12, (fused, "oldrhs", c)
            assert np.allclose(sc[:3], o.fr, rtol=1e-12, atol=1e-15), (fused, sc[:3], o.fr)
            assert abs(sc[3] - o.corrpx) <= 1e-11 * max(1.0, abs(o.corrpx))
            assert abs(sc[5] - mp1[0]) <= 1e-12 * max(1.0, abs(mp1[0]))
            ny_, nz_ = p.ny, p.nz                      # wall values of the mean column, as outstats reads them
            for k, ref in ((10, o.V[0, 0:5, 0, nz_].real), (15, o.V[0, ny_ - 2:ny_ + 3, 0, nz_].real),
                           (20, o.V[2, 0:5, 0, nz_].real), (25, o.V[2, ny_ - 2:ny_ + 3, 0, nz_].real)):
                assert np.allclose(sc[k:k + 5], ref, rtol=1e-11, atol=1e-13), (fused, k)
        o.meanpx, o.meanpz = mp1


@pytest.mark.parametrize("hook", ["coriolis", "am_f1", "am_butterfly", "am_f1_wide"])
def test_body_force_kernels_match_the_reference_hooks(emul, hook):
    """body_force_kernel / force_ghost_kernel (bodyforce_kernels.cu) with the masks channel_b200.dnsdata builds,
    against the oracle's restatement of body_forces/*/*.inc and of the ghost extension dnsdata.f90:616-629."""
    from channel_b200.dnsdata import Channel

    class Masks:                       # Channel's mask builders without a handle (no GPU here)
        _am_pieces = Channel._am_pieces
        config_am_f1 = Channel.config_am_f1
        config_am_butterfly = Channel.config_am_butterfly
        config_coriolis = Channel.config_coriolis

        def config_body_force_linear(self, A, my, mz, exclude_mean=False):
            self.got = dict(A=A, my=my, mz=mz, myz=None, ex=exclude_mean)

        def config_body_force_linear_yz(self, A, myz, exclude_mean=False):
            self.got = dict(A=A, my=None, mz=None, myz=myz, ex=exclude_mean)

    nx, ny, nz = 5, 32, 7
    p = ODnsIn(nx=nx, ny=ny, nz=nz, re=1000.0)
    o = Oracle(p)
    rng = np.random.default_rng(11)
    o.V[:] = rng.standard_normal(o.V.shape) + 1j * rng.standard_normal(o.V.shape)
    m = Masks(); m.p = p; m.y = o.y; m.nz = nz; m.ny = ny
    if hook == "coriolis":
        fn = coriolis_force(0.02, 3.0, 0.4); m.config_coriolis(0.02, 3.0, 0.4)
    elif hook == "am_f1":
        fn = am_f1_force(2000.0, 10.0); m.config_am_f1(2000.0, 10.0)
    elif hook == "am_f1_wide":         # iz_f = 13 > nz: every stored z-mode is inside the |iz| <= iz_f range
        fn = am_f1_force(500.0, 1000.0); m.config_am_f1(500.0, 1000.0)
    else:
        fn = am_butterfly_force(2000.0, 10.0); m.config_am_butterfly(2000.0, 10.0)
    F0 = 0.1 * (rng.standard_normal(o.V.shape) + 0j)       # entries outside the mask keep their previous value
    o.F = F0.copy()
    fn(o)
    g = m.got
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64) if a is not None else None
    A, my, mz, myz = c(np.asarray(g["A"]).reshape(9)), c(g["my"]), c(g["mz"]), c(g["myz"])
    V = np.ascontiguousarray(o.V); F = np.ascontiguousarray(F0.copy())
    d4 = np.ascontiguousarray(o.d4[2:ny + 1], dtype=np.float64)
    assert emul.chb_emul_body_force(nx, ny, nz, _dp(V.view(np.float64)), _dp(F.view(np.float64)), _dp(A), _dp(my), _dp(mz),
                                    _dp(myz), int(g["ex"]), 0, _dp(d4)) == 0
    assert np.array_equal(F, o.F)
    assert (F != F0).any() and (F == F0).any()
    # ghost extension at the start of buildrhs
    Fo = o.F
    Fo[:, 0:2] = 0
    Fo[:, 0] = -o._D(o.d4, Fo, 1) / o.d4[2, 0]
    Fo[:, ny + 1:ny + 3] = 0
    Fo[:, ny + 2] = -o._D(o.d4, Fo, ny - 1) / o.d4[ny, 4]
    assert emul.chb_emul_body_force(nx, ny, nz, _dp(V.view(np.float64)), _dp(F.view(np.float64)), _dp(A), _dp(my), _dp(mz),
                                    _dp(myz), int(g["ex"]), 1, _dp(d4)) == 0
    assert relerr(F, Fo) < 1e-13
