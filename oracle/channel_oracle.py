"""CPU oracle for the per-timestep hot path of davecats/channel.

TEST INFRASTRUCTURE ONLY.  This is a numpy FP64 restatement of the reference
algorithm (Fortran/FFTW/MPI) used as the checker by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
Nothing under channel_b200/ (the product) may import it.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path, and its toolchain (gfortran, MPI, FFTW 3.x - un-vendored, version
unpinned, README.md:44) is absent from this image, so the reference itself
cannot be run here.  The oracle is pinned only by analytic known answers
(tests/test_oracle_known_answers.py): polynomial exactness of the compact-FD
tables, banded solves against dense numpy.linalg.solve, the laminar Poiseuille
fixed point, discrete continuity and the viscous decay of a mean spanwise mode.

FFT arithmetic (FFTW 3.x in the reference, ffts.f90:70-75) is restated with
numpy.fft / scipy.fft (pocketfft), same unnormalised conventions:
    IFT  : complex backward (sign +)   ffts.f90:71,93-96
    FFT  : complex forward  (sign -)   ffts.f90:70,88-91
    RFT  : c2r, logical length 2*nxd (sign +), imag of DC/Nyquist ignored  ffts.f90:72-73
    HFT  : r2c, logical length 2*nxd (sign -)  ffts.f90:74-75

Array conventions (all C-order numpy):
    V[c, iy+1, ix, iz+nz]  complex128, c=0..2 (u,v,w), iy=-1..ny+1, ix=0..nx, iz=-nz..nz
    der tables d0,d1,d2,d4[iy+1, j+2], valid for iy=1..ny-1, j=-2..2
All citations are file:line into /root/reference.
"""
from __future__ import annotations

import dataclasses
import os
import numpy as np

try:  # multi-threaded pocketfft when available (bench cpu_baseline uses all cores)
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None

_WORKERS = 1


def set_fft_workers(n: int) -> None:
    global _WORKERS
    _WORKERS = max(1, int(n))


def _ifft_u(a, axis):
    """unnormalised complex backward transform (sign +): ffts.f90:71."""
    if _sfft is not None:
        return _sfft.ifft(a, axis=axis, norm="forward", workers=_WORKERS)
    return np.fft.ifft(a, axis=axis, norm="forward")


def _fft_u(a, axis):
    """unnormalised complex forward transform (sign -): ffts.f90:70."""
    if _sfft is not None:
        return _sfft.fft(a, axis=axis, norm="backward", workers=_WORKERS)
    return np.fft.fft(a, axis=axis, norm="backward")


def _irfft_u(a, n, axis):
    """unnormalised c2r of logical length n (sign +): ffts.f90:72-73."""
    if _sfft is not None:
        return _sfft.irfft(a, n=n, axis=axis, norm="forward", workers=_WORKERS)
    return np.fft.irfft(a, n=n, axis=axis, norm="forward")


def _rfft_u(a, axis):
    """unnormalised r2c (sign -): ffts.f90:74-75."""
    if _sfft is not None:
        return _sfft.rfft(a, axis=axis, norm="backward", workers=_WORKERS)
    return np.fft.rfft(a, axis=axis, norm="backward")


# --------------------------------------------------------------------------
# sizes, parameters                                   dnsdata.f90:98-125
# --------------------------------------------------------------------------
def fft_fit(n: int) -> bool:
    """ffts.f90:78-86: n is 2^k or 3*2^k."""
    j = n
    while j % 2 == 0:
        j >>= 1
    return j == 1 or j == 3


def padded_sizes(nx: int, nz: int):
    """dnsdata.f90:110-113 (useFFTfit on, header.h:23)."""
    nxd = 3 * (nx + 1) // 2
    nzd = 3 * nz
    while not fft_fit(nxd):
        nxd += 1
    while not fft_fit(nzd):
        nzd += 1
    return nxd, nzd


RK1_rai = (120.0 / 32.0, 2.0, 0.0)            # dnsdata.f90:70
RK2_rai = (120.0 / 8.0, 50.0 / 8.0, 34.0 / 8.0)   # dnsdata.f90:71
RK3_rai = (120.0 / 20.0, 90.0 / 20.0, 50.0 / 20.0)  # dnsdata.f90:72


@dataclasses.dataclass
class DnsIn:
    """dns.in as read by read_dnsin (dnsdata.f90:110-122). `re` is the value on
    line 3 of dns.in; ni = 1/re (dnsdata.f90:115)."""
    nx: int = 16
    ny: int = 64
    nz: int = 16
    alfa0: float = 0.5
    beta0: float = 1.0
    re: float = 12431.0
    a: float = 1.5
    ymin: float = 0.0
    ymax: float = 2.0
    CPI: bool = True
    CPI_type: int = 1
    gamma: float = 0.161436
    meanpx: float = 0.0
    meanpz: float = 0.0
    meanflowx: float = 0.0
    meanflowz: float = 0.0
    u0: float = 0.0
    uN: float = 0.0
    deltat: float = 0.0
    cflmax: float = 1.0
    time: float = 0.0


class Oracle:
    """State + hot path of MODULE dnsdata, serial (npx=npy=1)."""

    def __init__(self, p: DnsIn):
        self.p = p
        self.nx, self.ny, self.nz = p.nx, p.ny, p.nz
        self.nxd, self.nzd = padded_sizes(p.nx, p.nz)
        nx, ny, nz, nxd, nzd = self.nx, self.ny, self.nz, self.nxd, self.nzd
        self.alfa0, self.beta0 = p.alfa0, p.beta0
        self.ni = 1.0 / p.re                                   # dnsdata.f90:115
        self.meanpx, self.meanpz = p.meanpx, p.meanpz
        self.meanflowx, self.meanflowz = p.meanflowx, p.meanflowz
        self.CPI, self.CPI_type, self.gamma = p.CPI, p.CPI_type, p.gamma
        self.u0, self.uN = p.u0, p.uN
        self.deltat, self.cflmax, self.time = p.deltat, p.cflmax, p.time
        PI = 3.1415926535897932384626433832795028841971       # dnsdata.f90:26
        self.dx = PI / (p.alfa0 * nxd)                         # dnsdata.f90:124
        self.dz = 2.0 * PI / (p.beta0 * nzd)
        self.factor = 1.0 / (2.0 * nxd * nzd)
        # grid                                                 dnsdata.f90:153-155
        iy = np.arange(-1, ny + 2, dtype=np.float64)
        self.y = p.ymin + 0.5 * (p.ymax - p.ymin) * (
            np.tanh(p.a * (2.0 * iy / float(ny) - 1.0)) / np.tanh(p.a) + 1.0)
        self.dy = np.zeros(ny + 3)
        self.dy[2:ny + 1] = 0.5 * (self.y[3:ny + 2] - self.y[1:ny])   # iy=1..ny-1
        # wavenumbers                                          dnsdata.f90:156-158
        ix = np.arange(0, nx + 1)
        iz = np.arange(-nz, nz + 1)
        self.ialfa = 1j * (ix * p.alfa0)
        self.ibeta = 1j * (iz * p.beta0)
        self.k2 = (p.alfa0 * ix)[:, None] ** 2.0 + (p.beta0 * iz)[None, :] ** 2.0   # [ix, iz]
        self.izd = np.where(iz >= 0, iz, nzd + iz)             # dnsdata.f90:156
        # state
        self.V = np.zeros((3, ny + 3, nx + 1, 2 * nz + 1), np.complex128)
        self.oldrhs = np.zeros((2, ny + 3, nx + 1, 2 * nz + 1), np.complex128)  # [eta,d2v]; A.7: zero
        self.F = None          # body force (same shape as V) or None
        # convection-velocity diagnostic (#ifdef convvel, dnsdata.f90:84-89,141-143): off unless enable_convvel()
        self.convvel = False
        self.Voldz = None; self.uconv = None
        self.convvel_cnt = -1; self.compute_convvel = False
        self.cfl = 0.0
        self.fr = np.zeros(3)
        self.corrpx = 0.0
        self.corrpz = 0.0
        self.setup_derivatives()
        self.setup_boundary_conditions()

    # ----------------------------------------------------------------------
    # rbmat.f90:60-76 LUdecomp, rbmat.f90:201-215 LLUdiv (.bs.)
    # ----------------------------------------------------------------------
    @staticmethod
    def LUdecomp(A):
        HI = A.shape[0]
        for i in range(HI - 1, 0, -1):          # i = HI..2 (1-based)
            piv = 1.0 / A[i, i]
            A[i, i] = piv
            A[i, :i] = A[i, :i] * piv
            for k in range(0, i):
                piv = A[k, i]
                A[k, :i] = A[k, :i] - piv * A[i, :i]
        A[0, 0] = 1.0 / A[0, 0]

    @staticmethod
    def bs(A, b):
        HI = A.shape[0]
        x = np.zeros(HI)
        x[HI - 1] = b[HI - 1] * A[HI - 1, HI - 1]
        for i in range(HI - 2, -1, -1):
            x[i] = (b[i] - np.sum(A[i, i + 1:HI] * x[i + 1:HI])) * A[i, i]
        for i in range(1, HI):
            x[i] = x[i] - np.sum(A[i, :i] * x[:i])
        return x

    # ----------------------------------------------------------------------
    # setup_derivatives                                  dnsdata.f90:241-286
    # ----------------------------------------------------------------------
    def setup_derivatives(self):
        ny, y = self.ny, self.y
        Y = lambda i: y[i + 1]
        self.d0 = np.zeros((ny + 3, 5)); self.d1 = np.zeros((ny + 3, 5))
        self.d2 = np.zeros((ny + 3, 5)); self.d4 = np.zeros((ny + 3, 5))
        ii = np.arange(5, dtype=np.float64)
        for iy in range(1, ny):
            h = np.array([Y(iy - 2 + j) - Y(iy) for j in range(5)])
            M = np.power(h[None, :], (4.0 - ii)[:, None]); self.LUdecomp(M)       # :247
            t = np.zeros(5); t[0] = 24.0
            d4 = self.bs(M, t)                                                    # :249
            M = ((5.0 - ii) * (6.0 - ii) * (7.0 - ii) * (8.0 - ii))[:, None] * np.power(h[None, :], (4.0 - ii)[:, None])
            self.LUdecomp(M)                                                      # :250
            t = np.array([np.sum(d4 * np.power(h, 8.0 - i)) for i in range(5)])   # :251
            d0 = self.bs(M, t)                                                    # :252
            M = np.power(h[None, :], (4.0 - ii)[:, None]); self.LUdecomp(M)       # :253
            t = np.zeros(5)
            for i in range(3):
                t[i] = np.sum(d0 * (4.0 - i) * (3.0 - i) * np.power(h, 2.0 - i))  # :254
            d2 = self.bs(M, t)
            t = np.zeros(5)
            for i in range(4):
                t[i] = np.sum(d0 * (4.0 - i) * np.power(h, 3.0 - i))              # :256
            d1 = self.bs(M, t)
            self.d0[iy + 1], self.d1[iy + 1], self.d2[iy + 1], self.d4[iy + 1] = d0, d1, d2, d4

        def wall(nodes0, base):
            h = np.array([Y(nodes0 + j) - Y(base) for j in range(5)])
            M = np.power(h[None, :], (4.0 - ii)[:, None]); self.LUdecomp(M)
            t = np.zeros(5); t[3] = 1.0; a1 = self.bs(M, t)
            t = np.zeros(5); t[2] = 2.0; a2 = self.bs(M, t)
            return a1, a2
        self.d140, self.d240 = wall(-1, 0)             # :260-262
        self.d14m1, self.d24m1 = wall(-1, -1)          # :263-265
        self.d040 = np.zeros(5); self.d040[1] = 1.0    # :266  d040(-1)=1
        self.d14n, self.d24n = wall(ny - 3, ny)        # :269-271
        self.d14np1, self.d24np1 = wall(ny - 3, ny + 1)  # :272-274
        self.d04n = np.zeros(5); self.d04n[3] = 1.0    # :275  d04n(1)=1
        # D0mat(ny0:nyN+2,-2:2), rows iy=1..ny+1; rows ny,ny+1 zero (A.7)   :277,284
        D0mat = np.zeros((ny + 1, 5))
        D0mat[0:ny - 1] = self.d0[2:ny + 1]
        self.LU5decompStep(D0mat)
        self.D0mat = D0mat

    # ----------------------------------------------------------------------
    # setup_boundary_conditions                          dnsdata.f90:290-308
    # ----------------------------------------------------------------------
    def setup_boundary_conditions(self):
        ny = self.ny
        v0bc = self.d040.copy(); v0m1bc = self.d140.copy(); eta0bc = self.d040.copy()
        eta0m1bc = self.d4[1 + 1].copy()                                       # der(1)%d4
        v0bc[1:5] = v0bc[1:5] - v0bc[0] * v0m1bc[1:5] / v0m1bc[0]
        eta0bc[1:5] = eta0bc[1:5] - eta0bc[0] * eta0m1bc[1:5] / eta0m1bc[0]
        vnbc = self.d04n.copy(); vnp1bc = self.d14n.copy(); etanbc = self.d04n.copy()
        etanp1bc = self.d4[ny - 1 + 1].copy()                                  # der(ny-1)%d4
        vnbc[0:4] = vnbc[0:4] - vnbc[4] * vnp1bc[0:4] / vnp1bc[4]
        etanbc[0:4] = etanbc[0:4] - etanbc[4] * etanp1bc[0:4] / etanp1bc[4]
        self.v0bc, self.v0m1bc, self.eta0bc, self.eta0m1bc = v0bc, v0m1bc, eta0bc, eta0m1bc
        self.vnbc, self.vnp1bc, self.etanbc, self.etanp1bc = vnbc, vnp1bc, etanbc, etanp1bc

    # ----------------------------------------------------------------------
    # rbparmat_blocking.f90:20-100 (npy=1: first=last=.TRUE.)
    # A has shape (ny+1, 5[, ncol]); row index i <-> iy=i+1; band index j+2.
    # x has leading shape (ny+3,) ; x index i+2 <-> Fortran x(i), i=-2..ny <-> iy=i+1
    # ----------------------------------------------------------------------
    @staticmethod
    def LU5decompStep(A):
        HI1 = A.shape[0] - 1
        A[HI1 - 2, 3:5] = 0.0; A[HI1 - 3, 4] = 0.0                  # :29
        for i in range(HI1 - 2, -1, -1):                            # :35
            for k in (2, 1):
                piv = A[i, k + 2].copy()
                for j in (-1, -2):
                    A[i, j + k + 2] = A[i, j + k + 2] - piv * A[i + k, j + 2]
            piv = 1.0 / A[i, 2]
            A[i, 2] = piv
            A[i, 0] = A[i, 0] * piv
            A[i, 1] = A[i, 1] * piv
        A[0, 0:2] = 0.0; A[1, 0] = 0.0                               # :45

    @staticmethod
    def LeftLU5divStep1(A, x):
        """in place; x shape (ny+3, ...) indexed iy+1.  :57-77"""
        HI1 = A.shape[0] - 1
        for i in range(HI1 - 2, -1, -1):
            # Fortran x(i) <-> iy = i+1 <-> python index i+2
            x[i + 2] = (x[i + 2] - (A[i, 3] * x[i + 3] + A[i, 4] * x[i + 4])) * A[i, 2]

    @staticmethod
    def LeftLU5divStep2(A, b):
        """in place.  :82-100"""
        HI1 = A.shape[0] - 1
        for i in range(0, HI1 + 1):
            b[i + 2] = b[i + 2] - (A[i, 0] * b[i] + A[i, 1] * b[i + 1])

    # ----------------------------------------------------------------------
    # yintegr                                            dnsdata.f90:312-324
    # ----------------------------------------------------------------------
    def yintegr(self, f):
        """f indexed iy+1, real."""
        y = self.y
        II = 0.0
        for iy in range(1, self.ny, 2):                 # (ny0/2)*2+1 = 1 .. nyN step 2
            yp1 = y[iy + 2] - y[iy + 1]; ym1 = y[iy] - y[iy + 1]
            a1 = -1.0 / 3.0 * ym1 + 1.0 / 6.0 * yp1 + 1.0 / 6.0 * yp1 * yp1 / ym1
            a3 = +1.0 / 3.0 * yp1 - 1.0 / 6.0 * ym1 - 1.0 / 6.0 * ym1 * ym1 / yp1
            a2 = yp1 - ym1 - a1 - a3
            II = II + a1 * f[iy] + a2 * f[iy + 1] + a3 * f[iy + 2]
        return II

    # ----------------------------------------------------------------------
    # COMPLEXderiv (+ the LeftLU5divStep2 its callers add)   dnsdata.f90:339-373
    # ----------------------------------------------------------------------
    def COMPLEXderiv_full(self, f0):
        """f0 shape (ny+3, ...) -> compact first derivative, same shape."""
        ny = self.ny
        f1 = np.zeros_like(f0)
        bsh = (5,) + (1,) * (f0.ndim - 1)
        f1[1] = np.sum(self.d140.reshape(bsh) * f0[0:5], axis=0)            # f1(0)
        f1[0] = np.sum(self.d14m1.reshape(bsh) * f0[0:5], axis=0)           # f1(-1)
        f1[ny + 1] = np.sum(self.d14n.reshape(bsh) * f0[ny - 2:ny + 3], axis=0)     # f1(ny)
        f1[ny + 2] = np.sum(self.d14np1.reshape(bsh) * f0[ny - 2:ny + 3], axis=0)   # f1(ny+1)
        csh = (ny - 1,) + (1,) * (f0.ndim - 1)
        acc = 0
        for j in range(5):
            acc = acc + self.d1[2:ny + 1, j].reshape(csh) * f0[j:j + ny - 1]
        f1[2:ny + 1] = acc
        d0 = self.d0
        f1[2] = f1[2] - (d0[2, 1] * f1[1] + d0[2, 0] * f1[0])                # :361
        f1[3] = f1[3] - d0[3, 0] * f1[1]                                     # :362
        f1[ny] = f1[ny] - (d0[ny, 3] * f1[ny + 1] + d0[ny, 4] * f1[ny + 2])  # :365
        f1[ny - 1] = f1[ny - 1] - d0[ny - 1, 4] * f1[ny + 1]                 # :366
        A = self.D0mat.reshape(self.D0mat.shape + (1,) * (f0.ndim - 1))
        self.LeftLU5divStep1(A, f1)
        self.LeftLU5divStep2(A, f1)
        return f1

    # ----------------------------------------------------------------------
    # convolutions, all planes at once                   dnsdata.f90:487-602
    # returns VVd[6, ny+3, nx+1, nzd]  (= VVdz(1:nzd,1:nxB,1:6,slot) per plane)
    # ----------------------------------------------------------------------
    def convolutions(self, V, compute_cfl: bool):
        nx, ny, nz, nxd, nzd = self.nx, self.ny, self.nz, self.nxd, self.nzd
        Z = np.zeros((3, ny + 3, nx + 1, nzd), np.complex128)
        Z[..., 0:nz + 1] = V[..., nz:2 * nz + 1]           # :504
        Z[..., nzd - nz:nzd] = V[..., 0:nz]                 # :505
        Z = _ifft_u(Z, axis=-1)                             # :510  IFT
        if self.convvel and self.compute_convvel:           # :515-531, all planes of this sweep; then :546-549
            if self.convvel_cnt > -1:
                ix0 = np.arange(self.nx + 1, dtype=np.float64)[None, None, :, None]
                dtu = (Z - self.Voldz) / self.deltat
                ust = 0.5 * (Z + self.Voldz)
                with np.errstate(divide="ignore", invalid="ignore"):
                    cu = (np.conj(ust) * dtu).imag / (ix0 * self.alfa0 * (ust * np.conj(ust)).real)
                self.uconv[:, :, 1:, :] += cu[:, :, 1:, :]  # IF (ix0>0)
            self.Voldz = Z.copy()
            self.convvel_cnt += 1
            self.compute_convvel = False
        X = np.zeros((3, ny + 3, nxd + 1, nzd), np.complex128)
        X[:, :, 0:nx + 1, :] = Z                            # :533 zTOx, :535 zero pad
        R = _irfft_u(X, n=2 * nxd, axis=2)                  # :535  RFT -> [3, ny+3, 2nxd, nzd]
        del X, Z
        if compute_cfl:                                     # :552-556, planes 1..ny-1
            s = (np.abs(R[0, 2:ny + 1]) / self.dx
                 + np.abs(R[1, 2:ny + 1]) / self.dy[2:ny + 1, None, None]
                 + np.abs(R[2, 2:ny + 1]) / self.dz)
            self.cfl = max(self.cfl, float(s.max()))
        f = self.factor
        P = np.empty((6,) + R.shape[1:])
        P[3] = R[0] * R[1] * f                              # :581
        P[4] = R[1] * R[2] * f                              # :582
        P[5] = R[0] * R[2] * f                              # :583
        P[0:3] = R[0:3] * R[0:3] * f                        # :584
        del R
        H = _rfft_u(P, axis=2)[:, :, 0:nx + 1, :]           # :586 HFT, :588 xTOz keeps 0..nx
        del P
        return _fft_u(H, axis=-1)                           # :590 FFT

    # ----------------------------------------------------------------------
    # buildrhs                                           dnsdata.f90:611-673
    # Returns rhs[2, ny+3, nx+1, 2nz+1] (0 = eta-rhs, 1 = D2v-rhs; rows 1..ny-1 set)
    # and writes it into V(1:ny-1, :, :, 1:2) exactly like :667-671.
    # ----------------------------------------------------------------------
    def buildrhs(self, ODE, compute_cfl: bool):
        nx, ny, nz = self.nx, self.ny, self.nz
        V, ni, k2 = self.V, self.ni, self.k2
        deltat = self.deltat
        F = self.F
        if F is not None:                                   # :616-629 ghost extension of F
            F[:, 0:2] = 0
            F[:, 0] = -self._D(self.d4, F, 1) / self.d4[2, 0]
            F[:, ny + 1:ny + 3] = 0
            F[:, ny + 2] = -self._D(self.d4, F, ny - 1) / self.d4[ny, 4]
        VVd = self.convolutions(V, compute_cfl)
        P = VVd[..., self.izd]                              # truncation in z (DD macro :609)
        del VVd
        ia = self.ialfa[None, :, None]
        ib = self.ibeta[None, None, :]
        k2b = k2[None]
        DD = lambda d, k: self._Dall(d, P[k])               # [ny-1, nx+1, 2nz+1]
        DD0_6 = DD(self.d0, 5); DD1_6 = DD(self.d1, 5)
        rhsu = -ia * DD(self.d0, 0) - DD(self.d1, 3) - ib * DD0_6           # :638
        rhsv = -ia * DD(self.d0, 3) - DD(self.d1, 1) - ib * DD(self.d0, 4)  # :639
        rhsw = -ia * DD0_6 - DD(self.d1, 4) - ib * DD(self.d0, 2)           # :640
        expl = (ia * (ia * DD(self.d1, 0) + DD(self.d2, 3) + ib * DD1_6)
                + ib * (ia * DD1_6 + DD(self.d2, 4) + ib * DD(self.d1, 2)) - k2b * rhsv)  # :641-642
        if F is not None:                                   # :644
            expl = expl - k2b * self._Dall(self.d0, F[1]) - ia * self._Dall(self.d1, F[0]) - ib * self._Dall(self.d1, F[2])
        sl = slice(2, ny + 1)
        rhs = np.zeros((2, ny + 3, nx + 1, 2 * nz + 1), np.complex128)
        old = self.oldrhs
        # D2v equation                                                       :647
        unkn = self._Dall(self.d2, V[1]) - k2b * self._Dall(self.d0, V[1])
        OSc = [ni * (self.d4[sl, j, None, None] - 2.0 * k2b * self.d2[sl, j, None, None]
                     + k2b * k2b * self.d0[sl, j, None, None]) for j in range(5)]         # :476
        impl = sum(OSc[j] * V[1, j:j + ny - 1] for j in range(5))
        rhs[1, sl] = ODE[0] * unkn / deltat + impl + ODE[1] * expl - ODE[2] * old[1, sl]  # :486
        old[1, sl] = expl
        # eta equation, general modes                                        :656-661
        expl = ib * rhsu - ia * rhsw
        if F is not None:
            expl = expl + ib * self._Dall(self.d0, F[0]) - ia * self._Dall(self.d0, F[2])
        unkn = ib * self._Dall(self.d0, V[0]) - ia * self._Dall(self.d0, V[2])
        SQc = [ni * (self.d2[sl, j, None, None] - k2b * self.d0[sl, j, None, None]) for j in range(5)]  # :477
        impl = sum(SQc[j] * (ib * V[0, j:j + ny - 1] - ia * V[2, j:j + ny - 1]) for j in range(5))
        eta_rhs = ODE[0] * unkn / deltat + impl + ODE[1] * expl - ODE[2] * old[0, sl]
        # mean mode (ix=0,iz=0)                                              :648-654
        expl00 = (rhsu[:, 0, nz].real + self.meanpx) + 1j * (rhsw[:, 0, nz].real + self.meanpz)
        if F is not None:
            expl00 = expl00 + self._rD(self.d0, F, 0, 2)
        unkn00 = self._rD(self.d0, V, 0, 2)
        impl00 = ni * self._rD(self.d2, V, 0, 2)
        eta00 = ODE[0] * unkn00 / deltat + impl00 + ODE[1] * expl00 - ODE[2] * old[0, sl, 0, nz]
        old[0, sl] = expl
        old[0, sl, 0, nz] = expl00
        rhs[0, sl] = eta_rhs
        rhs[0, sl, 0, nz] = eta00
        V[0, sl] = rhs[0, sl]                               # :669
        V[1, sl] = rhs[1, sl]
        return rhs

    def _Dall(self, d, f):
        """sum_j d(iy,j) f(iy+j) for iy=1..ny-1; f shape (ny+3, ...)."""
        ny = self.ny
        acc = 0
        for j in range(5):
            acc = acc + d[2:ny + 1, j].reshape((ny - 1,) + (1,) * (f.ndim - 1)) * f[j:j + ny - 1]
        return acc

    def _D(self, d, F, iy):
        """stencil at a single iy for all comps: F shape (3, ny+3, ...)."""
        acc = 0
        for j in range(5):
            acc = acc + d[iy + 1, j] * F[:, iy - 2 + j + 1]
        return acc

    def _rD(self, d, f, g, k):
        """rD0/rD2 macro (dnsdata.f90:326-328) on column (0,0): real(f_g) + i real(f_k)."""
        nz = self.nz
        return (self._Dall(d, f[g][:, 0, nz].real) + 1j * self._Dall(d, f[k][:, 0, nz].real))

    # ----------------------------------------------------------------------
    # linsolve (blocking variant, incl. inline vetaTOuvw)   linsolve_blocking.inc:3-107
    # ----------------------------------------------------------------------
    def linsolve(self, lam: float, block: int = 4096):
        nx, ny, nz = self.nx, self.ny, self.nz
        V = self.V
        nzt = 2 * nz + 1
        M = (nx + 1) * nzt
        Vf = V.reshape(3, ny + 3, M)
        k2f = self.k2.reshape(M)
        iaf = np.repeat(self.ialfa, nzt)
        ibf = np.tile(self.ibeta, nx + 1)
        m00 = 0 * nzt + nz
        sl = slice(2, ny + 1)
        for c0 in range(0, M, block):
            c1 = min(M, c0 + block)
            k2 = k2f[c0:c1][None, :]
            n = c1 - c0
            D2vmat = np.zeros((ny + 1, 5, n)); etamat = np.zeros((ny + 1, 5, n))
            for j in range(5):
                d0 = self.d0[sl, j, None]; d2 = self.d2[sl, j, None]; d4 = self.d4[sl, j, None]
                OS = self.ni * (d4 - 2.0 * k2 * d2 + k2 * k2 * d0)
                SQ = self.ni * (d2 - k2 * d0)
                D2vmat[0:ny - 1, j] = lam * (d2 - k2 * d0) - OS            # :12
                etamat[0:ny - 1, j] = lam * d0 - SQ                        # :13
            # wall data (A.6): only (0,0) carries non-zero BCs
            bc0_v = np.zeros(n, complex); bc0_vy = np.zeros(n, complex); bc0_eta = np.zeros(n, complex)
            bcn_v = np.zeros(n, complex); bcn_vy = np.zeros(n, complex); bcn_eta = np.zeros(n, complex)
            if c0 <= m00 < c1:
                bc0_eta[m00 - c0] = complex(self.u0, 0.0)                   # :17
                bcn_eta[m00 - c0] = complex(self.uN, 0.0)                   # :31
            v = Vf[1, :, c0:c1]; eta = Vf[0, :, c0:c1]
            v0bc, v0m1bc, eta0bc, eta0m1bc = self.v0bc, self.v0m1bc, self.eta0bc, self.eta0m1bc
            vnbc, vnp1bc, etanbc, etanp1bc = self.vnbc, self.vnp1bc, self.etanbc, self.etanp1bc
            bc0_v = bc0_v - v0bc[0] * bc0_vy / v0m1bc[0]                    # :21
            self.applybc_0(D2vmat, v0bc, v0m1bc)                            # :22
            v[2] = v[2] - D2vmat[0, 0] * bc0_vy / v0m1bc[0] - D2vmat[0, 1] * bc0_v / v0bc[1]   # :23
            v[3] = v[3] - D2vmat[1, 0] * bc0_v / v0bc[1]                    # :24
            self.applybc_0(etamat, eta0bc, eta0m1bc)                        # :25
            eta[2] = eta[2] - etamat[0, 1] * bc0_eta / eta0bc[1]            # :26
            eta[3] = eta[3] - etamat[1, 0] * bc0_eta / eta0bc[1]            # :27
            bcn_v = bcn_v - vnbc[4] * bcn_vy / vnp1bc[4]                    # :35
            self.applybc_n(D2vmat, vnbc, vnp1bc)                            # :36
            v[ny] = v[ny] - D2vmat[ny - 2, 4] * bcn_vy / vnp1bc[4] - D2vmat[ny - 2, 3] * bcn_v / vnbc[3]  # :37
            v[ny - 1] = v[ny - 1] - D2vmat[ny - 3, 4] * bcn_v / vnbc[3]     # :38
            self.applybc_n(etamat, etanbc, etanp1bc)                        # :39
            eta[ny] = eta[ny] - etamat[ny - 2, 3] * bcn_eta / etanbc[3]     # :40
            eta[ny - 1] = eta[ny - 1] - etamat[ny - 3, 4] * bcn_eta / etanbc[3]  # :41
            self.LU5decompStep(D2vmat); self.LU5decompStep(etamat)          # :43
            self.LeftLU5divStep1(D2vmat, v)                                 # :44
            self.LeftLU5divStep1(etamat, eta)                               # :45
            self.LeftLU5divStep2(D2vmat, v)                                 # :48
            self.LeftLU5divStep2(etamat, eta)                               # :49
            v[1] = (bc0_v - (v[2] * v0bc[2] + v[3] * v0bc[3] + v[4] * v0bc[4])) / v0bc[1]                        # :51
            v[0] = (bc0_vy - (v[1] * v0m1bc[1] + v[2] * v0m1bc[2] + v[3] * v0m1bc[3] + v[4] * v0m1bc[4])) / v0m1bc[0]  # :52
            eta[1] = (bc0_eta - (eta[2] * eta0bc[2] + eta[3] * eta0bc[3] + eta[4] * eta0bc[4])) / eta0bc[1]      # :53
            eta[0] = -(eta[1] * eta0m1bc[1] + eta[2] * eta0m1bc[2] + eta[3] * eta0m1bc[3] + eta[4] * eta0m1bc[4]) / eta0m1bc[0]  # :54
            v[ny + 1] = (bcn_v - (v[ny - 2] * vnbc[0] + v[ny - 1] * vnbc[1] + v[ny] * vnbc[2])) / vnbc[3]        # :57
            v[ny + 2] = (bcn_vy - (v[ny - 2] * vnp1bc[0] + v[ny - 1] * vnp1bc[1] + v[ny] * vnp1bc[2] + v[ny + 1] * vnp1bc[3])) / vnp1bc[4]  # :58
            eta[ny + 1] = (bcn_eta - (eta[ny - 2] * etanbc[0] + eta[ny - 1] * etanbc[1] + eta[ny] * etanbc[2])) / etanbc[3]   # :59
            eta[ny + 2] = -(eta[ny - 2] * etanp1bc[0] + eta[ny - 1] * etanp1bc[1] + eta[ny] * etanp1bc[2] + eta[ny + 1] * etanp1bc[3]) / etanp1bc[4]  # :60
            has00 = c0 <= m00 < c1
            if has00:
                eta00 = eta[:, m00 - c0].copy()
                eta00mat = etamat[:, :, m00 - c0].copy()
            # vetaTOuvw for all modes of the block (the (0,0) column is overwritten below)  :99-103
            vy = self.COMPLEXderiv_full(v)
            k2s = np.where(k2 == 0.0, 1.0, k2)
            ia = iaf[c0:c1][None, :]; ib = ibf[c0:c1][None, :]
            temp = (ia * vy - ib * eta) / k2s
            w = (ib * vy + ia * eta) / k2s
            Vf[2, :, c0:c1] = w
            Vf[0, :, c0:c1] = temp
            if has00:                                                        # :62-97
                i0 = m00 - c0
                U = eta00.real.copy(); W = eta00.imag.copy()                 # :63-64
                ucor = np.zeros(ny + 3); ucor[2:ny + 1] = 1.0                # :65
                self.LeftLU5divStep1(eta00mat, ucor)                         # :67
                self.LeftLU5divStep2(eta00mat, ucor)                         # :68
                ucor[1] = -(ucor[2] * eta0bc[2] + ucor[3] * eta0bc[3] + ucor[4] * eta0bc[4]) / eta0bc[1]                 # :70
                ucor[0] = -(ucor[1] * eta0m1bc[1] + ucor[2] * eta0m1bc[2] + ucor[3] * eta0m1bc[3] + ucor[4] * eta0m1bc[4]) / eta0m1bc[0]  # :71
                ucor[ny + 1] = -(ucor[ny - 2] * etanbc[0] + ucor[ny - 1] * etanbc[1] + ucor[ny] * etanbc[2]) / etanbc[3]   # :74
                ucor[ny + 2] = -(ucor[ny - 2] * etanp1bc[0] + ucor[ny - 1] * etanp1bc[1] + ucor[ny] * etanp1bc[2] + ucor[ny + 1] * etanp1bc[3]) / etanp1bc[4]  # :75
                self.fr = np.array([self.yintegr(U), self.yintegr(W), self.yintegr(ucor)])   # :77-78
                if abs(self.meanflowx) > 1.0e-7 and not self.CPI:            # :79-82
                    self.corrpx = (self.meanflowx - self.fr[0]) / self.fr[2]
                    U = U + self.corrpx * ucor
                if abs(self.meanflowz) > 1.0e-7 and not self.CPI:            # :83-86
                    self.corrpz = (self.meanflowz - self.fr[1]) / self.fr[2]
                    W = W + self.corrpz * ucor
                if self.CPI:                                                 # :87-97
                    if self.CPI_type == 0:
                        self.meanpx = (1 - self.gamma) * 6 * self.ni / self.fr[0]
                    elif self.CPI_type == 1:
                        self.meanpx = (1.5 / self.gamma) * self.fr[0] * self.ni
                    else:
                        raise ValueError("Wrong selection of CPI_Type")
                Vf[0, :, m00] = U
                Vf[2, :, m00] = W

    @staticmethod
    def applybc_0(EQ, bc0, bc0m1):
        """dnsdata.f90:458-464; EQ rows: index 0 <-> iy=1."""
        e = EQ[0, 0].copy()
        for j in range(1, 5):
            EQ[0, j] = EQ[0, j] - e * bc0m1[j] / bc0m1[0]
        e = EQ[0, 1].copy()
        for j in range(2, 5):
            EQ[0, j] = EQ[0, j] - e * bc0[j] / bc0[1]
        e = EQ[1, 0].copy()
        for j in range(1, 4):
            EQ[1, j] = EQ[1, j] - e * bc0[j + 1] / bc0[1]

    @staticmethod
    def applybc_n(EQ, bcn, bcnp1):
        """dnsdata.f90:466-472; row ny-1 <-> index ny-2."""
        r1 = EQ.shape[0] - 3      # iy = ny-1
        r2 = EQ.shape[0] - 4      # iy = ny-2
        e = EQ[r1, 4].copy()
        for j in range(0, 4):
            EQ[r1, j] = EQ[r1, j] - e * bcnp1[j] / bcnp1[4]
        e = EQ[r1, 3].copy()
        for j in range(0, 3):
            EQ[r1, j] = EQ[r1, j] - e * bcn[j] / bcn[3]
        e = EQ[r2, 4].copy()
        for j in range(1, 4):
            EQ[r2, j] = EQ[r2, j] - e * bcn[j - 1] / bcn[3]

    # ----------------------------------------------------------------------
    # driver pieces                                      channel.f90:95-179
    # ----------------------------------------------------------------------
    def cfl_prepass(self):
        """channel.f90:96-115: CFL over planes 1..ny-1, flow rate, CPI meanpx."""
        if self.deltat == 0:
            self.deltat = 1.0
        self.convolutions(self.V, True)
        nz = self.nz
        self.fr[0] = self.yintegr(self.V[0, :, 0, nz].real)
        self.fr[1] = self.yintegr(self.V[2, :, 0, nz].real)
        if self.CPI:
            if self.CPI_type == 0:
                self.meanpx = (1 - self.gamma) * 6 * self.ni / self.fr[0]
            elif self.CPI_type == 1:
                self.meanpx = (1.5 / self.gamma) * self.fr[0] * self.ni

    def outstats(self):
        """dnsdata.f90:853-880: returns the Runtimedata line (11 columns)."""
        ny, nz = self.ny, self.nz
        if self.convvel:
            self.compute_convvel = True                     # :858-860
        runtime_global = self.cfl
        self.cfl = 0.0
        if self.cflmax > 0:
            self.deltat = self.cflmax / runtime_global
        U = self.V[0, :, 0, nz].real; W = self.V[2, :, 0, nz].real
        dudy0 = np.sum(self.d140 * U[0:5]); dwdy0 = np.sum(self.d140 * W[0:5])
        dudyN = -np.sum(self.d14n * U[ny - 2:ny + 3]); dwdyN = -np.sum(self.d14n * W[ny - 2:ny + 3])
        return np.array([self.time, dudy0, dudyN, dwdy0, dwdyN,
                         self.fr[0] + self.corrpx * self.fr[2], self.meanpx + self.corrpx,
                         self.fr[1] + self.corrpz * self.fr[2], self.meanpz + self.corrpz,
                         runtime_global * self.deltat, self.deltat])

    def enable_convvel(self):
        """#define convvel: Voldz / uconv(1:nzd, 1:nxB, iy, 1:3), here [3, ny+3, nx+1, nzd] (dnsdata.f90:141-143)."""
        self.convvel = True
        shape = (3, self.ny + 3, self.nx + 1, self.nzd)
        self.Voldz = np.zeros(shape, np.complex128); self.uconv = np.zeros(shape)
        self.convvel_cnt = -1; self.compute_convvel = False

    def convvel_file_bytes(self):
        """what outstats writes at the dt_field cadence (dnsdata.f90:908-913, save_convvel_file :792-816):
        uconv / convvel_cnt as float64 [iV][iy+1][ix][iz_d], no header; then uconv = 0, convvel_cnt = 0."""
        out = (self.uconv / self.convvel_cnt).astype(np.float64).tobytes()
        self.uconv[:] = 0; self.convvel_cnt = 0
        return out

    def set_body_force(self, fn):
        """fn(oracle) fills self.F from self.V (body_forces/*.inc hooks)."""
        self._body_force = fn

    def step(self):
        """one RK3 step, channel.f90:118-167."""
        for k, RK in enumerate((RK1_rai, RK2_rai, RK3_rai)):
            self.time = self.time + 2.0 / RK[0] * self.deltat
            if getattr(self, "_body_force", None) is not None:
                self._body_force(self)
            self.buildrhs(RK, k == 2)
            self.linsolve(RK[0] / self.deltat)
        return self.outstats()


def coriolis_force(omega2: float, kz_cutoff: float, y_threshold_bot: float):
    """body_forces/coriolis/coriolis.inc:4-41."""
    def fn(o: Oracle):
        if o.F is None:
            o.F = np.zeros_like(o.V)
        y_threshold_top = o.p.ymax - y_threshold_bot
        iz_thr = min(o.nz, int(np.floor(kz_cutoff / o.beta0)))
        ymask = (o.y <= y_threshold_bot) | (o.y >= y_threshold_top)
        zs = slice(o.nz - iz_thr, o.nz + iz_thr + 1)
        for c in (1, 2):
            pc = c % 2 + 1
            zeichen = 2 * pc - 3
            o.F[pc - 1][ymask, :, zs] = zeichen * omega2 * o.V[c - 1][ymask, :, zs]
    return fn


def _am_masks(o: Oracle, lambdaz_f: float):
    """shared pieces of the am hooks: iz_f (config_body_force, am_f1.inc:7) and y+ of every node (:20)."""
    iz_f = int(np.floor((2.0 * np.pi / lambdaz_f) / (o.beta0 / 1000.0) + 0.5))   # NINT
    yp = np.where(o.y > 1, o.p.ymax - o.y, o.y) * 1000.0
    return iz_f, yp


def am_f1_force(lambdaz_f: float = 500.0, amp: float = 1000.0):
    """body_forces/am_f1/am_f1.inc:13-28 (parameters am_pardec.inc:1): F = -amp V where
    lambda_z+ > 2.3 (y+)^2, |iz| <= iz_f, except the mean mode."""
    def fn(o: Oracle):
        if o.F is None:
            o.F = np.zeros_like(o.V)
        iz_f, yp = _am_masks(o, lambdaz_f)
        for iz in range(-min(iz_f, o.nz), min(iz_f, o.nz) + 1):
            lzp = 1e10 if iz == 0 else 2 * np.pi / (o.beta0 * abs(iz)) * 1000
            ym = lzp > 2.3 * yp ** 2
            for ix in range(o.nx + 1):
                if ix == 0 and iz == 0:
                    continue
                o.F[:, ym, ix, iz + o.nz] = -amp * o.V[:, ym, ix, iz + o.nz]
    return fn


def am_butterfly_force(lambdaz_f: float = 500.0, amp: float = 1000.0):
    """body_forces/am_butterfly/am_butterfly.inc:11-29: F = -amp V for (|iz| <= iz_f, y+ <= 60, not the mean
    mode) and for (|iz| > iz_f, y+ > 60)."""
    def fn(o: Oracle):
        if o.F is None:
            o.F = np.zeros_like(o.V)
        iz_f, yp = _am_masks(o, lambdaz_f)
        for iz in range(-o.nz, o.nz + 1):
            ym = (yp <= 60) if abs(iz) <= iz_f else (yp > 60)
            for ix in range(o.nx + 1):
                if abs(iz) <= iz_f and ix == 0 and iz == 0:
                    continue
                o.F[:, ym, ix, iz + o.nz] = -amp * o.V[:, ym, ix, iz + o.nz]
    return fn

