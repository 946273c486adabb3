"""CPU tests that pin the oracle (the reference ships no tests or golden vectors for this path,
SURVEY.md section 4 / 8c): analytic known answers, agreement of the two independent
restatements (numpy all-planes-at-once vs C plane-by-plane), and the committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

from channel_b200.fields import perturbed_laminar
from oracle.channel_oracle import DnsIn, Oracle, RK1_rai, RK2_rai, RK3_rai, coriolis_force, padded_sizes
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
