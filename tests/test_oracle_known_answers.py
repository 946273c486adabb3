"""CPU tests that pin the oracle (the reference ships no tests or golden vectors for this path,
SURVEY.md section 4 / 8c): analytic known answers, agreement of the two independent
restatements (numpy all-planes-at-once vs C plane-by-plane), and the committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

from channel_b200.fields import perturbed_laminar
from oracle.channel_oracle import (DnsIn, Oracle, RK1_rai, RK2_rai, RK3_rai, am_butterfly_force, am_f1_force,
                                   coriolis_force, padded_sizes)
from oracle.c_oracle import COracle, fft_lines

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_padded_sizes_follow_fftfit():
    # dnsdata.f90:110-113 + ffts.f90:78-86; values of SURVEY.md section 8 table
    assert padded_sizes(16, 16) == (32, 48)
    assert padded_sizes(191, 189) == (384, 768)
    assert padded_sizes(511, 511) == (768, 1536)
    assert padded_sizes(1023, 1023) == (1536, 3072)
    assert padded_sizes(383, 383) == (768, 1536)


@pytest.mark.parametrize("ny", [16, 64])
def test_compact_fd_tables_are_exact_on_polynomials(ny):
    """d0 f' = d1 f and d0 f'' = d2 f, d0 f'''' = d4 f hold exactly for polynomials up to degree 4
    (the defining property of the compact scheme, dnsdata.f90:246-258)."""
    o = Oracle(DnsIn(nx=4, ny=ny, nz=4))
    y = o.y
    for deg in range(5):
        f = y ** deg
        f1 = deg * y ** max(deg - 1, 0) if deg >= 1 else 0 * y
        f2 = deg * (deg - 1) * y ** max(deg - 2, 0) if deg >= 2 else 0 * y
        f4 = 24.0 * np.ones_like(y) if deg == 4 else 0 * y
        for iy in range(1, ny):
            s = slice(iy - 1, iy + 4)
            scale = max(1.0, np.abs(o.d1[iy + 1]).max())
            assert abs(o.d1[iy + 1] @ f[s] - o.d0[iy + 1] @ f1[s]) < 1e-9 * scale
            scale = max(1.0, np.abs(o.d2[iy + 1]).max())
            assert abs(o.d2[iy + 1] @ f[s] - o.d0[iy + 1] @ f2[s]) < 1e-8 * scale
            scale = max(1.0, np.abs(o.d4[iy + 1]).max())
            assert abs(o.d4[iy + 1] @ f[s] - o.d0[iy + 1] @ f4[s]) < 1e-7 * scale
    # one-sided wall stencils: first derivative exact on quartics
    f = y ** 4; f1 = 4 * y ** 3
    assert abs(o.d140 @ f[0:5] - f1[1]) < 1e-9
    assert abs(o.d14n @ f[ny - 2:ny + 3] - f1[ny + 1]) < 1e-9


def test_compact_derivative_of_smooth_function():
    o = Oracle(DnsIn(nx=4, ny=96, nz=4))
    f = np.sin(1.3 * o.y) + 0.2j * np.cos(0.7 * o.y)
    d = o.COMPLEXderiv_full(f)
    exact = 1.3 * np.cos(1.3 * o.y) - 0.14j * np.sin(0.7 * o.y)
    assert np.abs(d - exact).max() < 5e-6


def test_banded_ul_solver_matches_dense_solve():
    """LU5decompStep + LeftLU5divStep1/2 (rbparmat_blocking.f90:20-100) solve A x = b."""
    rng = np.random.default_rng(3)
    n = 40                                     # rows iy = 1..ny-1 with ny-1 = n
    A = rng.standard_normal((n + 2, 5)) + np.array([0, 0, 8, 0, 0])
    A[n:] = 0.0                                # halo rows
    dense = np.zeros((n, n))
    for i in range(n):
        for j in range(-2, 3):
            if 0 <= i + j < n:
                dense[i, i + j] = A[i, j + 2]
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = np.zeros(n + 4, complex); x[2:n + 2] = b
    LU = A.copy()
    Oracle.LU5decompStep(LU)
    Oracle.LeftLU5divStep1(LU, x)
    Oracle.LeftLU5divStep2(LU, x)
    assert rel(x[2:n + 2], np.linalg.solve(dense, b)) < 1e-12


def test_fft_conventions():
    """IFT sign +, FFT sign -, both unnormalised (ffts.f90:70-75); C oracle's own FFT vs numpy."""
    rng = np.random.default_rng(0)
    for n in (8, 12, 32, 48, 96, 384, 768, 1536):
        x = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
        k = np.arange(n)
        dft = np.exp(-2j * np.pi * np.outer(k, k) / n)
        if n <= 96:
            assert rel(fft_lines(x, -1), x @ dft.T) < 1e-13
            assert rel(fft_lines(x, +1), x @ dft.conj().T) < 1e-13
        assert rel(fft_lines(x, -1), np.fft.fft(x, axis=1)) < 1e-14 * np.log2(n) * 4
        assert rel(fft_lines(x, +1), np.fft.ifft(x, axis=1, norm="forward")) < 1e-14 * np.log2(n) * 4


def test_convolution_is_the_dealiased_product():
    """For a field with a single pair of modes the products are known in closed form:
    u = 2 cos(a x) -> uu = 2 + 2 cos(2 a x): spectral uu(0)=2, uu(2)=1."""
    o = Oracle(DnsIn(nx=8, ny=8, nz=4))
    V = np.zeros_like(o.V)
    V[0, :, 1, o.nz] = 1.0                     # u_hat(ix=1, iz=0) = 1  ->  u = 2 cos(alfa0 x)
    P = o.convolutions(V, False)[..., o.izd]
    uu = P[0]
    assert np.allclose(uu[:, 0, o.nz], 2.0, atol=1e-13)
    assert np.allclose(uu[:, 2, o.nz], 1.0, atol=1e-13)
    uu2 = uu.copy(); uu2[:, 0, o.nz] = 0; uu2[:, 2, o.nz] = 0
    assert np.abs(uu2).max() < 1e-13


def test_laminar_poiseuille_is_a_fixed_point_and_continuity_holds():
    p = DnsIn(nx=8, ny=32, nz=6, re=1000.0, CPI=True, CPI_type=1, gamma=1.0, deltat=1e-2, cflmax=0.0)
    o = Oracle(p)
    o.V[0, :, 0, p.nz] = 1.5 * o.y * (2 - o.y)
    V0 = o.V.copy()
    o.cfl_prepass(); o.outstats()
    for _ in range(3):
        line = o.step()
    assert np.abs(o.V - V0).max() < 1e-12
    assert abs(line[1] - 3.0) < 1e-9 and abs(line[5] - 2.0) < 1e-12
    # perturbed: discrete continuity  i alfa u + D0^-1 D1 v + i beta w = 0 for every mode but (0,0)
    o = Oracle(DnsIn(nx=8, ny=32, nz=6, re=1000.0, deltat=1e-3, cflmax=0.0))
    o.V[:] = perturbed_laminar(8, 32, 6, 0.5, 1.0, eps=1e-2)
    o.cfl_prepass(); o.outstats(); o.step()
    vy = o.COMPLEXderiv_full(o.V[1])
    div = o.ialfa[None, :, None] * o.V[0] + vy + o.ibeta[None, None, :] * o.V[2]
    div[:, 0, o.nz] = 0
    assert np.abs(div).max() < 1e-13 * np.abs(o.V).max() * 100


def test_mean_spanwise_mode_decays_viscously():
    """W(y,t) = sin(pi y / 2) exp(-ni (pi/2)^2 t) solves W_t = ni W_yy with W=0 at the walls."""
    ny = 64
    p = DnsIn(nx=4, ny=ny, nz=4, re=50.0, CPI=False, deltat=2e-3, cflmax=0.0)
    o = Oracle(p)
    o.V[2, :, 0, p.nz] = np.sin(0.5 * np.pi * o.y)
    o.cfl_prepass(); o.outstats()
    nsteps = 50
    for _ in range(nsteps):
        o.step()
    t = nsteps * p.deltat
    exact = np.sin(0.5 * np.pi * o.y) * np.exp(-(1.0 / p.re) * (0.5 * np.pi) ** 2 * t)
    assert np.abs(o.V[2, 1:ny + 2, 0, p.nz].real - exact[1:ny + 2]).max() < 2e-6


@pytest.mark.parametrize("nx,ny,nz,couette", [(16, 64, 16, False), (7, 16, 5, False), (9, 20, 6, True)])
def test_c_and_numpy_restatements_agree(nx, ny, nz, couette):
    kw = dict(CPI=False, u0=-1.0, uN=1.0) if couette else {}
    p = DnsIn(nx=nx, ny=ny, nz=nz, re=3000.0, deltat=0.0, cflmax=1.0, **kw)
    o = Oracle(p); c = COracle(p)
    for name in COracle.TABLES:
        a = np.asarray(getattr(o, name)).reshape(-1); b = c.table(name).reshape(-1)
        assert rel(b, a) < 1e-13, name
    V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, eps=1e-2, couette=couette)
    o.V[:] = V0; c.set_V(V0)
    if couette:
        o.set_body_force(coriolis_force(0.02, 9999999.0, 1.0)); c.set_coriolis(0.02, 9999999.0, 1.0)
    o.cfl_prepass(); c.cfl_prepass()
    assert np.allclose(o.outstats(), c.outstats(), rtol=1e-12, atol=1e-13)
    for RK, last in ((RK1_rai, False), (RK2_rai, False), (RK3_rai, True)):   # substep by substep
        if couette:
            o._body_force(o); c.lib.co_set_body_force(c.h)
        o.buildrhs(RK, last); c.buildrhs(RK, last)
        Vc = c.get_V()
        assert rel(Vc[0, 2:ny + 1], o.V[0, 2:ny + 1]) < 1e-12 and rel(Vc[1, 2:ny + 1], o.V[1, 2:ny + 1]) < 1e-12
        lam = RK[0] / o.deltat
        o.linsolve(lam); c.linsolve(lam)
        Vc = c.get_V()
        for k in range(3):
            assert rel(Vc[k], o.V[k]) < 1e-12
    for _ in range(3):
        lo = o.step(); lc = c.step()
        assert np.allclose(lo, lc, rtol=1e-9, atol=1e-11)
    Vc = c.get_V()
    for k in range(3):
        assert rel(Vc[k], o.V[k]) < 1e-11
    c.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_oracles_reproduce_golden_fixtures(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    name = os.path.splitext(os.path.basename(path))[0]
    pk, fk, cor, nsteps = mg.CASES[name]
    g = np.load(path)
    p = DnsIn(**pk)
    assert np.array_equal(perturbed_laminar(p.nx, p.ny, p.nz, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, **fk), g["V0"])
    for cls in (Oracle, COracle):
        o = cls(p)
        if cls is Oracle:
            o.V[:] = g["V0"]
            if cor:
                o.set_body_force(coriolis_force(0.02, 9999999.0, 1.0))
            for t in ("d0", "d1", "d2", "d4", "D0mat", "y", "v0bc", "eta0m1bc", "vnbc", "etanp1bc"):
                assert np.array_equal(getattr(o, t), g[t]), t
        else:
            o.set_V(g["V0"])
            if cor:
                o.set_coriolis(0.02, 9999999.0, 1.0)
        o.cfl_prepass()
        lines = [o.outstats()]
        for i in range(nsteps):
            lines.append(o.step())
            if i == 0:
                V1 = o.V.copy() if cls is Oracle else o.get_V()
        Vend = o.V if cls is Oracle else o.get_V()
        tol = 0.0 if cls is Oracle else 1e-11
        for k in range(3):
            assert rel(V1[k], g["V1"][k]) <= tol and rel(Vend[k], g["Vend"][k]) <= tol * 10
        assert np.allclose(np.array(lines), g["lines"], rtol=max(tol * 100, 1e-15), atol=1e-11 if tol else 0)


def test_fd_tables_match_the_references_matlab_formulation():
    """The reference carries a second, independently written implementation of its compact-FD tables:
    matlab-interface/base/compute_derivatives.m:18-79 (grid: init_dns.m:24, walls==2), which solves the same
    5x5 moment systems with MATLAB's pivoted `M\\t` instead of the Fortran's unpivoted LUdecomp / LLUdiv
    (rbmat.f90:60-76,201-215).  Restated here with numpy.linalg.solve and compared with the oracle's
    (dnsdata.f90:241-286 order of operations) and the C++ host tables: a cross-check of the restatement
    against the reference's own alternative formulation (the wall rows of d1/d2 are the d14*/d24* stencils)."""
    from channel_b200.dnsdata import Tables
    ny, a, ymin, ymax = 48, 1.5, 0.0, 2.0
    o = Oracle(DnsIn(nx=2, ny=ny, nz=2))
    y = 0.5 * (1 + np.tanh(a * (2 * np.arange(-1, ny + 2) / ny - 1)) / np.tanh(a)) * (ymax - ymin) + ymin   # init_dns.m:24
    assert np.allclose(y, o.y, rtol=0, atol=1e-15)
    d = {k: np.zeros((ny + 3, 5)) for k in ("d0", "d1", "d2", "d4")}
    for iY in range(1, ny):                                     # compute_derivatives.m:18-54
        iy = iY + 1                                             # 0-based row of node iY
        dyv = y[iy - 2:iy + 3] - y[iy]
        M = np.array([[dyv[j] ** (4 - i) for j in range(5)] for i in range(5)])
        t = np.zeros(5); t[0] = 24
        d4 = np.linalg.solve(M, t)
        M0 = np.array([[(5 - i) * (6 - i) * (7 - i) * (8 - i) * dyv[j] ** (4 - i) for j in range(5)] for i in range(5)])
        t = np.array([np.sum(d4 * dyv ** (8 - i)) for i in range(5)])
        d0 = np.linalg.solve(M0, t)
        t = np.zeros(5)
        for i in range(3):
            t[i] = np.sum(d0 * (4 - i) * (3 - i) * dyv ** (2 - i))
        d2 = np.linalg.solve(M, t)
        t = np.zeros(5)
        for i in range(4):
            t[i] = np.sum(d0 * (4 - i) * dyv ** (3 - i))
        d1 = np.linalg.solve(M, t)
        d["d4"][iy], d["d0"][iy], d["d2"][iy], d["d1"][iy] = d4, d0, d2, d1
    def wall(node, base):                                       # compute_derivatives.m:55-90
        dyv = y[base:base + 5] - y[node]
        M = np.array([[dyv[j] ** (4 - i) for j in range(5)] for i in range(5)])
        t1 = np.zeros(5); t1[3] = 1
        t2 = np.zeros(5); t2[2] = 2
        return np.linalg.solve(M, t1), np.linalg.solve(M, t2)
    ref_wall = {"d140": wall(1, 0)[0], "d240": wall(1, 0)[1], "d14m1": wall(0, 0)[0], "d24m1": wall(0, 0)[1],
                "d14n": wall(ny + 1, ny - 2)[0], "d24n": wall(ny + 1, ny - 2)[1],
                "d14np1": wall(ny + 2, ny - 2)[0], "d24np1": wall(ny + 2, ny - 2)[1]}
    host = Tables(ny, a, ymin, ymax)
    rel = lambda x, r: np.abs(x - r).max() / np.abs(r).max()
    for k in ("d0", "d1", "d2", "d4"):
        for iy in range(2, ny + 1):
            assert rel(getattr(o, k)[iy], d[k][iy]) < 1e-9, (k, iy, rel(getattr(o, k)[iy], d[k][iy]))
            assert rel(getattr(host, k)[iy - 2], d[k][iy]) < 1e-9, ("host", k, iy)
    for k, r in ref_wall.items():
        assert rel(np.asarray(getattr(o, k)), r) < 1e-9, (k, rel(np.asarray(getattr(o, k)), r))
        assert rel(np.asarray(getattr(host, k)), r) < 1e-9, ("host", k)


def test_transform_conventions_match_the_references_plane_ift():
    """matlab-interface/base/plane_ift.m is the reference's own statement of the spectral -> physical
    transform of a plane: z-modes (-nz..nz) stored as (0..nz, -nz..-1) in a zero-padded line of nzd, ifft along
    z, 'symmetric' ifft of logical length 2 nxd along x, times 2 nzd nxd (i.e. unnormalised, sign +).  The
    nonlinear term of the oracle must equal: physical fields by plane_ift -> pointwise products ->
    the adjoint-convention forward transforms written as explicit DFT sums (sign -, dnsdata.f90:124 factor
    1/(2 nxd nzd)) -> modes 0..nx, -nz..nz (mpi_transpose.f90:99-106, izd dnsdata.f90:156)."""
    nx, ny, nz = 3, 8, 2
    o = Oracle(DnsIn(nx=nx, ny=ny, nz=nz))
    nxd, nzd = o.nxd, o.nzd
    V = perturbed_laminar(nx, ny, nz, o.alfa0, o.beta0, eps=0.3, seed=3)
    ref = o.convolutions(V, False)[..., o.izd]                       # [6, ny+3, nx+1, 2nz+1]

    def plane_ift(A):                                                # A[ix 0..nx, iz+nz] -> real [2nxd, nzd]
        out = np.zeros((2 * nxd, nzd), complex)
        out[0:nx + 1, nzd - nz:nzd] = A[:, 0:nz]                     # plane_ift.m:11-13
        out[0:nx + 1, 0:nz + 1] = A[:, nz:2 * nz + 1]
        out[0:nx + 1] = np.fft.ifft(out[0:nx + 1], axis=1)           # :16-18 (MATLAB ifft = 1/N sum X e^{+i...})
        herm = np.zeros((2 * nxd, nzd), complex)                     # :21 'symmetric': the lower half is the conjugate mirror
        herm[0:nx + 1] = out[0:nx + 1]
        herm[0] = herm[0].real
        for k in range(1, nx + 1):
            herm[2 * nxd - k] = np.conj(herm[k])
        phys = np.fft.ifft(herm, axis=0) * (2 * nzd * nxd)
        assert np.abs(phys.imag).max() < 1e-12 * max(1.0, np.abs(phys.real).max())
        return phys.real

    x = np.arange(2 * nxd); z = np.arange(nzd)
    Ex = np.exp(-2j * np.pi * np.outer(np.arange(nx + 1), x) / (2 * nxd))                 # forward in x, modes 0..nx
    kz = np.arange(-nz, nz + 1)
    Ez = np.exp(-2j * np.pi * np.outer(z, kz) / nzd)                                      # forward in z, modes -nz..nz
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]                              # uu vv ww uv vw uw (dnsdata.f90:581-584)
    for iy in (0, 3, ny + 2):
        phys = [plane_ift(V[c, iy]) for c in range(3)]
        for k, (a, b) in enumerate(pairs):
            P = Ex @ (phys[a] * phys[b] * o.factor) @ Ez
            assert np.abs(P - ref[k, iy]).max() <= 1e-13 * max(1e-30, np.abs(ref[k]).max()), (iy, k)


@pytest.mark.parametrize("hook", ["am_f1", "am_butterfly"])
def test_c_and_numpy_restatements_agree_on_the_am_hooks(hook):
    """body_forces/am_f1/am_f1.inc and am_butterfly/am_butterfly.inc in both restatements (numpy: masks over whole
    arrays; C: the hooks' own loop nests), three RK3 steps."""
    nx, ny, nz = 9, 32, 7
    p = DnsIn(nx=nx, ny=ny, nz=nz, re=1500.0, deltat=2e-3, cflmax=0.0)
    o = Oracle(p); c = COracle(p)
    V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, eps=1e-2)
    o.V[:] = V0; c.set_V(V0)
    o.set_body_force((am_f1_force if hook == "am_f1" else am_butterfly_force)(2000.0, 10.0)); c.set_am(hook, 2000.0, 10.0)
    o.cfl_prepass(); c.cfl_prepass()
    assert np.allclose(o.outstats(), c.outstats(), rtol=1e-12, atol=1e-13)
    for _ in range(3):
        lo = o.step(); lc = c.step()
        assert np.allclose(lo, lc, rtol=1e-9, atol=1e-11)
    Vc = c.get_V()
    for k in range(3):
        assert rel(Vc[k], o.V[k]) < 1e-11
    assert np.abs(o.F).max() > 0
