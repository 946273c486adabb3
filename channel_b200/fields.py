"""Synthetic initial fields (host side, numpy).

The reference generates its own initial field with a zero-amplitude perturbation
(dnsdata.f90:706-718: U = 1.5 y (2-y) in mode (0,0)).  For parity and throughput
runs we use the perturbed-laminar field described in SURVEY.md section 8(d).

Layout: V[c, iy+1, ix, iz+nz] complex128 (the device layout).
"""
import numpy as np


def grid_y(ny, a=1.5, ymin=0.0, ymax=2.0):
    """dnsdata.f90:153 tanh grid, iy=-1..ny+1."""
    iy = np.arange(-1, ny + 2, dtype=np.float64)
    return ymin + 0.5 * (ymax - ymin) * (np.tanh(a * (2.0 * iy / float(ny) - 1.0)) / np.tanh(a) + 1.0)


def perturbed_laminar(nx, ny, nz, alfa0, beta0, a=1.5, ymin=0.0, ymax=2.0,
                      eps=1e-3, seed=20261017, couette=False):
    """Laminar profile in (0,0) plus eps*(1+k2)^-1 * g(y) * (xi1 + i xi2) in every other mode,
    g(y) = (y(2-y))^2 (zero with zero slope at both walls), Hermitian on the ix=0 line,
    v(0,0) = 0."""
    y = grid_y(ny, a, ymin, ymax)
    rng = np.random.default_rng(seed)
    V = np.zeros((3, ny + 3, nx + 1, 2 * nz + 1), np.complex128)
    ix = np.arange(nx + 1); iz = np.arange(-nz, nz + 1)
    k2 = (alfa0 * ix)[:, None] ** 2 + (beta0 * iz)[None, :] ** 2
    amp = eps / (1.0 + k2)
    g = (y * (2.0 - y)) ** 2
    # smooth-in-y random content: a few random y-shapes so that derivatives stay O(1)
    nshape = 3
    for c in range(3):
        for s in range(nshape):
            xi = rng.standard_normal((nx + 1, 2 * nz + 1)) + 1j * rng.standard_normal((nx + 1, 2 * nz + 1))
            shape = g * np.cos(0.5 * np.pi * s * y + 0.3 * c)
            V[c] += shape[:, None, None] * (amp * xi)[None]
    # Hermitian symmetry on ix=0: V(-iz,0) = conj V(iz,0)
    V[:, :, 0, :nz] = np.conj(V[:, :, 0, :nz:-1])
    V[:, :, 0, nz] = 0.0
    V[0, :, 0, nz] = (y - 1.0) if couette else 1.5 * y * (2.0 - y)
    return V


def perturbed_laminar_slab(out, nx, ny, nz, alfa0, beta0, nx0, nxB, a=1.5, ymin=0.0, ymax=2.0,
                           eps=1e-3, seed=20261017, couette=False):
    """The same field as perturbed_laminar(), for the x-slab ix = nx0..nx0+nxB-1 only, written
    into `out` in the Fortran / Dati.cart.out layout V(iy,iz,ix,c) = C-order [c][ixl][iz+nz][iy+1]
    (dnsdata.f90:132).  One BLAS call per component; no full-size temporaries."""
    assert out.shape == (3, nxB, 2 * nz + 1, ny + 3) and out.dtype == np.complex128
    y = grid_y(ny, a, ymin, ymax)
    rng = np.random.default_rng(seed)
    ix = np.arange(nx + 1); iz = np.arange(-nz, nz + 1)
    k2 = (alfa0 * ix)[:, None] ** 2 + (beta0 * iz)[None, :] ** 2
    amp = eps / (1.0 + k2)
    g = (y * (2.0 - y)) ** 2
    nshape = 3
    nzt = 2 * nz + 1
    for c in range(3):
        coef = np.empty((nxB, nzt, nshape), np.complex128)
        shapes = np.empty((nshape, ny + 3), np.complex128)
        for s in range(nshape):
            xi = rng.standard_normal((nx + 1, nzt)) + 1j * rng.standard_normal((nx + 1, nzt))
            coef[:, :, s] = (amp * xi)[nx0:nx0 + nxB]
            shapes[s] = g * np.cos(0.5 * np.pi * s * y + 0.3 * c)
        if nx0 == 0:   # Hermitian symmetry on ix=0 and the mean mode, as in perturbed_laminar()
            coef[0, :nz, :] = np.conj(coef[0, :nz:-1, :])
            coef[0, nz, :] = 0.0
        np.matmul(coef.reshape(nxB * nzt, nshape), shapes, out=out[c].reshape(nxB * nzt, ny + 3))
    if nx0 == 0:
        out[0, 0, nz, :] = (y - 1.0) if couette else 1.5 * y * (2.0 - y)
    return out
