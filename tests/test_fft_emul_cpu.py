"""The FFT-pass kernels of the nonlinear term (xpass3_kernels.cu, zpass3_kernels.cu) run on the CPU.

tests/host_emul/fft_emul.cpp compiles the kernel SOURCE with g++ and runs each thread block on OS threads
(cta_emul.hpp: __syncthreads = barrier, shared memory = a buffer).  For a few z-lines of each specialised
size (nxd = 384, 768, 1536) the result - zero-padded c2r, CFL, six products * factor, r2c, modes 0..nx in the
tiled work-buffer layout - is compared with numpy's FFT (ffts.f90:72-75 conventions, dnsdata.f90:535-588).
The GPU tests (tests/test_fft3_gpu.py) check the same kernels on the device; this one needs no GPU and is
where a new kernel variant is proven before it is measured."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "fft_emul.cpp")
BUILD = os.path.join(HERE, "host_emul", "_build")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emul():
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libfft_emul.so")
    csrc = os.path.join(HERE, "..", "channel_b200", "csrc")
    deps = [SRC, os.path.join(HERE, "host_emul", "cta_emul.hpp")] + [
        os.path.join(csrc, f) for f in ("xpass3_kernels.cu", "zpass3_kernels.cu", "fft_regs.cuh", "fft_device.cuh", "chb_internal.h", "transpose_index.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-pthread",
                               "-fvisibility=hidden", "-Wl,-Bsymbolic", "-I" + CUDA_INC, "-o", so, SRC])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.chb_emul_xpass.argtypes = [C.c_int] * 6 + [C.c_double] * 2 + [C.c_int, dp, dp, dp, C.c_int, dp, C.c_int]
    lib.chb_emul_xpass.restype = C.c_int
    lib.chb_emul_zpass.argtypes = [C.c_int] * 8 + [dp, dp]
    lib.chb_emul_zpass.restype = C.c_int
    lib.chb_emul_convolutions_multi.argtypes = [C.c_int, dp, dp]
    lib.chb_emul_convolutions_multi.restype = C.c_int
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def xpass_reference(A, nx, nxd, nzd, alfa0, beta0, dy, ny, compute_cfl):
    """A: [3][np][nzB][nx+1] -> products [6][np][nzB][nx+1], cfl"""
    _, npl, nzB, _ = A.shape
    X = np.zeros((3, npl, nzB, nxd + 1), complex)
    X[..., :nx + 1] = A                                               # zero padding in x, dnsdata.f90:535
    R = np.fft.irfft(X, n=2 * nxd, axis=-1) * (2 * nxd)               # RFT, unnormalised, sign +
    factor = 1.0 / (2.0 * nxd * nzd)                                  # dnsdata.f90:124
    dx = np.pi / (alfa0 * nxd); dz = 2 * np.pi / (beta0 * nzd)
    cfl = 0.0
    if compute_cfl:
        for pli in range(npl):
            iy = pli - 1
            if 1 <= iy <= ny - 1:
                s = np.abs(R[0, pli]) / dx + np.abs(R[1, pli]) / dy[iy + 1] + np.abs(R[2, pli]) / dz
                cfl = max(cfl, float(s.max()))
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]          # dnsdata.f90:581-584
    P = np.stack([R[a] * R[b] * factor for a, b in pairs])
    return np.fft.rfft(P, axis=-1)[..., :nx + 1], cfl                 # HFT, sign -, keep 0..nx


@pytest.mark.parametrize("nx,nxd,tw,variant", [(255, 384, 3, 0), (255, 384, 0, 0), (511, 768, 3, 0), (1023, 1536, 3, 0),
                                               (300, 768, 0, 0), (1023, 1536, 3, 1), (703, 1536, 2, 1),
                                               (511, 768, 3, 1), (300, 768, 0, 1),
                                               (1023, 1536, 3, 2), (703, 1536, 2, 3), (511, 768, 3, 2), (300, 768, 0, 3)])
def test_xpass_kernel_on_cpu_threads(emul, nx, nxd, tw, variant):
    ny, nzd, nzB, npl = 6, 6, 2, 3            # planes iy = -1, 0, 1: the CFL expression sees iy = 1 only
    rng = np.random.default_rng(nx + tw)
    A = rng.standard_normal((3, npl, nzB, nx + 1)) + 1j * rng.standard_normal((3, npl, nzB, nx + 1))
    dy = np.linspace(0.01, 0.03, ny + 3)
    ref, cfl_ref = xpass_reference(A, nx, nxd, nzd, 0.5, 1.0, dy, ny, True)
    Ar = np.ascontiguousarray(A)
    Br = np.zeros((6, npl, (nx + 1) >> tw, nzB, 1 << tw), complex)
    cfl = C.c_double()
    rc = emul.chb_emul_xpass(nx, ny, nzB, npl, nxd, nzd, 0.5, 1.0, tw, _dp(Ar.view(np.float64)), _dp(Br.view(np.float64)),
                             _dp(dy), 1, C.byref(cfl), variant)
    assert rc == 0
    got = np.transpose(Br, (0, 1, 3, 2, 4)).reshape(6, npl, nzB, nx + 1)      # [p][plane][z row][x tile][x in tile] -> x
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-13 * scale, np.abs(got - ref).max() / scale
    assert abs(cfl.value - cfl_ref) <= 1e-13 * cfl_ref


def _tiled(a, tw):
    """[..., z row, x] -> work-buffer layout [..., x tile, z row, x in tile] (transpose_index.h)"""
    *lead, nz_, nx_ = a.shape
    return np.ascontiguousarray(np.moveaxis(a.reshape(*lead, nz_, nx_ >> tw, 1 << tw), -2, -3))


@pytest.mark.parametrize("nz,nzd,lpc", [(255, 768, 4), (255, 768, 2), (255, 768, 8), (511, 1536, 4), (511, 1536, 2),
                                        (511, 1536, 8), (1023, 3072, 2), (1023, 3072, 4), (300, 768, 4),
                                        (1023, 3072, 2 + 1024), (511, 1536, 2 + 1024),       # + 1024: 128 threads per line
                                        (511, 1536, 2 + 2048)])                              # + 2048: 96 threads per line
def test_zpass_kernels_on_cpu_threads(emul, nz, nzd, lpc):
    """zfwd4: zero-pad in z + backward FFT + zTOx pack (dnsdata.f90:504-510, ffts.f90:71, mpi_transpose.f90:64-71);
    zbwd4: xTOz unpack + forward FFT + truncation through izd (ffts.f90:70, dnsdata.f90:609), incl. the TMA /
    cp.async staging replaced by plain copies."""
    nxB, npl = 8, 3
    nzt = 2 * nz + 1
    lpc_code, lpc = lpc, lpc & 1023
    rng = np.random.default_rng(nz + lpc)
    V = rng.standard_normal((3, npl, nxB, nzt)) + 1j * rng.standard_normal((3, npl, nxB, nzt))
    Z = np.zeros((3, npl, nxB, nzd), complex)
    Z[..., 0:nz + 1] = V[..., nz:]
    Z[..., nzd - nz:] = V[..., :nz]
    ref = np.fft.ifft(Z, axis=-1) * nzd                                # [3][np][x][z]
    for twa, mode in ((-1, 1), ({2: 1, 4: 2, 8: 3}[lpc], 1), (-1, 2), ({2: 1, 4: 2, 8: 3}[lpc], 2)):   # mode 2 = DIRECT stage A
        out = np.zeros((3, npl, nzd, nxB), complex)
        assert emul.chb_emul_zpass(mode, nxB, nz, nzd, npl, lpc_code, 3, twa, _dp(np.ascontiguousarray(V).view(np.float64)),
                                   _dp(out.view(np.float64))) == 0
        want = np.swapaxes(ref, -1, -2)                                # [3][np][z][x]
        want = want if twa < 0 else _tiled(want, twa).reshape(out.shape)
        assert np.abs(out - want).max() <= 1e-13 * np.abs(ref).max(), (twa, np.abs(out - want).max())
    # backward pass: products in the tiled buffer -> spectral, truncated
    B = rng.standard_normal((6, npl, nzd, nxB)) + 1j * rng.standard_normal((6, npl, nzd, nxB))
    F = np.fft.fft(np.swapaxes(B, -1, -2), axis=-1)                    # [6][np][x][k]
    refP = np.concatenate([F[..., nzd - nz:], F[..., :nz + 1]], axis=-1)
    for tw in sorted({0, {2: 1, 4: 2, 8: 3}[lpc], 3}):
        P = np.zeros((6, npl, nxB, nzt), complex)
        assert emul.chb_emul_zpass(0, nxB, nz, nzd, npl, lpc_code, tw, -1, _dp(_tiled(B, tw).view(np.float64)),
                                   _dp(P.view(np.float64))) == 0
        assert np.abs(P - refP).max() <= 1e-13 * np.abs(refP).max(), (tw, np.abs(P - refP).max())


@pytest.mark.parametrize("P", [1, 2, 8])
@pytest.mark.parametrize("persist", [False, True])
def test_nonlinear_term_on_emulated_ranks(emul, P, persist, monkeypatch):
    """zfwd -> (direct-store zTOx) -> xpass -> (direct-store xTOz) -> zbwd for one plane on P emulated ranks, each
    storing straight into the owners' buffers the way the GPUs do over NVLink (mpi_transpose.f90:50-117,214-215):
    the per-rank products must equal the single-domain numpy result, for every P (transpose invariance)."""
    if persist:        # the persistent x-pass: inputs of the next line prefetched into shared memory
        monkeypatch.setenv("CHB_EMUL_XPERSIST", "1")
    else:
        monkeypatch.delenv("CHB_EMUL_XPERSIST", raising=False)
    nx, nz, nxd, nzd = 255, 255, 384, 768
    nzt = 2 * nz + 1
    rng = np.random.default_rng(17)
    V = rng.standard_normal((3, nx + 1, nzt)) + 1j * rng.standard_normal((3, nx + 1, nzt))
    Z = np.zeros((3, nx + 1, nzd), complex)
    Z[..., :nz + 1] = V[..., nz:]; Z[..., nzd - nz:] = V[..., :nz]
    Z = np.fft.ifft(Z, axis=-1) * nzd
    X = np.zeros((3, nxd + 1, nzd), complex); X[:, :nx + 1] = Z
    R = np.fft.irfft(X, n=2 * nxd, axis=1) * (2 * nxd)
    f = 1.0 / (2.0 * nxd * nzd)
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    Pp = np.stack([R[a] * R[b] * f for a, b in pairs])
    H = np.fft.fft(np.fft.rfft(Pp, axis=1)[:, :nx + 1], axis=-1)
    ref = np.concatenate([H[..., nzd - nz:], H[..., :nz + 1]], axis=-1)            # [6][nx+1][2nz+1]
    nxB = (nx + 1) // P
    Vr = np.ascontiguousarray(np.stack([V[:, r * nxB:(r + 1) * nxB] for r in range(P)])[:, :, None])   # [P][3][1][nxB][nzt]
    out = np.zeros((P, 6, 1, nxB, nzt), complex)
    assert emul.chb_emul_convolutions_multi(P, _dp(Vr.view(np.float64)), _dp(out.view(np.float64))) == 0
    got = np.concatenate([out[r, :, 0] for r in range(P)], axis=1)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max(), np.abs(got - ref).max() / np.abs(ref).max()
