"""Parity of the CUDA hot path (through the C ABI) with the CPU oracle, FP64.

Tolerances are the ones BASELINE.json states: single-step RHS and fields 1e-12 relative,
fields 1e-9 after 10 steps, Runtimedata columns 2..9 to 1e-8 over the run.  "Relative" is
norm-wise per field (max|diff| / max|ref|), the way the reference's own compare_fields.py
reports differences.
"""
import numpy as np
import pytest

from channel_b200 import RK1_rai, RK2_rai, RK3_rai
from tests.helpers import make_pair, relerr

pytestmark = pytest.mark.gpu

GRIDS = [(16, 64, 16), (31, 48, 21), (7, 16, 5), (47, 40, 32)]


@pytest.mark.parametrize("nx,ny,nz", GRIDS)
def test_products_rhs_and_solve_one_substep(nx, ny, nz):
    p, o, ch, V0 = make_pair(nx, ny, nz)
    ch.cfl_prepass(); o.cfl_prepass()
    s = ch.get_step_scalars()
    assert abs(s["cfl"] - o.cfl) <= 1e-13 * o.cfl
    assert np.allclose(s["fr"][:2], o.fr[:2], rtol=1e-13, atol=1e-15)
    assert abs(s["meanpx"] - o.meanpx) <= 1e-13 * abs(o.meanpx)
    o.cfl = 0.0
    for RK, last in ((RK1_rai, False), (RK2_rai, False), (RK3_rai, True)):
        # products of the nonlinear term (VVdz after the forward z-FFT, truncated through izd)
        Pref = o.convolutions(o.V, False)[..., o.izd]
        ch.buildrhs(RK, last)
        Pgpu = ch.download_products()
        for k in range(6):
            assert relerr(Pgpu[k], Pref[k]) < 1e-12, ("product", k)
        rhs_ref = o.buildrhs(RK, last)
        rhs_gpu = ch.download_rhs()
        sl = slice(2, ny + 1)
        assert relerr(rhs_gpu[0, sl], rhs_ref[0, sl]) < 1e-12
        assert relerr(rhs_gpu[1, sl], rhs_ref[1, sl]) < 1e-12
        lam = RK[0] / p.deltat
        o.linsolve(lam); ch.linsolve(lam)
        Vg = ch.download_V()
        for c in range(3):
            assert relerr(Vg[c], o.V[c]) < 1e-12, ("field", c)
    s = ch.get_step_scalars()
    assert abs(s["cfl"] - o.cfl) <= 1e-12 * o.cfl
    assert np.allclose(s["fr"], o.fr, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("nx,ny,nz", [(16, 64, 16), (31, 48, 21)])
def test_ten_steps_and_runtimedata(nx, ny, nz):
    p, o, ch, V0 = make_pair(nx, ny, nz, deltat=0.0, cflmax=1.0, re=3000.0)
    ch.cfl_prepass(); o.cfl_prepass()
    l_g = ch.outstats(); l_o = o.outstats()
    assert np.allclose(l_g, l_o, rtol=1e-10, atol=1e-12)
    for i in range(10):
        l_o = o.step(); l_g = ch.step()
        assert np.allclose(l_g[1:9], l_o[1:9], rtol=1e-8, atol=1e-10), (i, l_g, l_o)
        assert np.allclose(l_g[[0, 9, 10]], l_o[[0, 9, 10]], rtol=1e-9)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-9


def test_hundred_steps_runtimedata_config1():
    """BASELINE config 1: nx,ny,nz=16,64,16, 100 RK3 steps; Runtimedata cols 2-9 to 1e-8."""
    p, o, ch, V0 = make_pair(16, 64, 16, deltat=0.0, cflmax=1.0, re=4000.0, eps=3e-2)
    ch.cfl_prepass(); o.cfl_prepass()
    ch.outstats(); o.outstats()
    for i in range(100):
        l_o = o.step(); l_g = ch.step()
        assert np.allclose(l_g[1:9], l_o[1:9], rtol=1e-8, atol=1e-10), (i, l_g, l_o)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-8


def test_laminar_poiseuille_fixed_point():
    """U = 1.5 y (2-y) with CPI type 1, gamma=1 is a steady state (SURVEY 8c known answer)."""
    from channel_b200 import Channel, DnsIn
    p = DnsIn(nx=16, ny=64, nz=16, re=1000.0, CPI=True, CPI_type=1, gamma=1.0, deltat=1e-2, cflmax=0.0)
    ch = Channel(p)
    V = np.zeros(ch.field_shape(), complex)
    y = ch.y
    V[0, :, 0, 16] = 1.5 * y * (2 - y)
    ch.upload_V(V)
    ch.cfl_prepass(); ch.outstats()
    for _ in range(5):
        line = ch.step()
    Vg = ch.download_V()
    assert np.abs(Vg - V).max() < 1e-12
    assert abs(line[1] - 3.0) < 1e-9 and abs(line[2] - 3.0) < 1e-9      # wall shear dU/dy = 3
    assert abs(line[5] - 2.0) < 1e-12                                    # flow rate


def test_couette_coriolis_body_force():
    """Couette walls (u0=-1,uN=1) + the coriolis hook (body_forces/coriolis/coriolis.inc)."""
    from oracle.channel_oracle import coriolis_force
    p, o, ch, V0 = make_pair(15, 32, 10, deltat=2e-3, cflmax=0.0, re=1500.0, couette=True,
                             CPI=False, u0=-1.0, uN=1.0)
    o.set_body_force(coriolis_force(0.02, 9999999.0, 1.0))
    ch.config_coriolis(0.02, 9999999.0, 1.0)
    ch.cfl_prepass(); o.cfl_prepass(); ch.outstats(); o.outstats()
    for i in range(3):
        l_o = o.step(); l_g = ch.step()
        assert np.allclose(l_g[1:9], l_o[1:9], rtol=1e-9, atol=1e-11), (i, l_g, l_o)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-11
    assert relerr(ch.download_F(), o.F) < 1e-12


def test_fortran_layout_roundtrip():
    """chb_upload_V / chb_download_V use the Fortran layout V(iy,iz,ix,c) of Dati.cart.out."""
    p, o, ch, V0 = make_pair(16, 64, 16)
    Vf = np.ascontiguousarray(np.transpose(V0, (0, 2, 3, 1)))     # [c][ix][iz][iy]
    ch.upload_V_fortran(Vf)
    assert np.array_equal(ch.download_V(), V0)
    assert np.array_equal(ch.download_V_fortran(), Vf)


@pytest.mark.parametrize("nx,ny,nz", [(2, 9, 1), (3, 12, 2), (6, 8, 1), (5, 8, 4)])
def test_minimal_grids(nx, ny, nz):
    """Smallest sizes the library accepts (ny = 8 is the minimum: two wall rows of each kind plus the
    interior; nz = 1 gives nzd = 3; odd ny; fewer rows than one checkpoint block of the banded solve)."""
    p, o, ch, V0 = make_pair(nx, ny, nz, deltat=1e-3, cflmax=0.0, re=500.0)
    ch.cfl_prepass(); o.cfl_prepass()
    assert np.allclose(ch.outstats(), o.outstats(), rtol=1e-11, atol=1e-13)
    for i in range(2):
        lo = o.step(); lg = ch.step()
        assert np.allclose(lg[1:9], lo[1:9], rtol=1e-9, atol=1e-11), (i, lg, lo)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-11, (c, relerr(Vg[c], o.V[c]))
    ch.close()


def test_rejected_sizes_report_errors():
    """Sizes outside the library's domain fail loudly at chb_create (no fallback): ny < 8, odd nxd."""
    from channel_b200 import Channel, DnsIn, _lib
    with pytest.raises(_lib.ChannelB200Error, match="ny>=8"):
        Channel(DnsIn(nx=4, ny=7, nz=4))
    with pytest.raises(_lib.ChannelB200Error, match="nxd must be even"):
        Channel(DnsIn(nx=1, ny=8, nz=1))           # nxd = 3(nx+1)/2 = 3


@pytest.mark.parametrize("hook", ["am_f1", "am_butterfly"])
def test_am_body_forces(hook):
    """body_forces/am_f1/am_f1.inc and am_butterfly/am_butterfly.inc: F = -amp V inside a region of the (y, z-mode)
    plane that is not a product of ranges (chb_set_body_force_linear_yz)."""
    from oracle.channel_oracle import am_butterfly_force, am_f1_force
    p, o, ch, V0 = make_pair(15, 32, 10, deltat=2e-3, cflmax=0.0, re=1500.0)
    if hook == "am_f1":
        o.set_body_force(am_f1_force(2000.0, 10.0)); ch.config_am_f1(2000.0, 10.0)
    else:
        o.set_body_force(am_butterfly_force(2000.0, 10.0)); ch.config_am_butterfly(2000.0, 10.0)
    ch.cfl_prepass(); o.cfl_prepass(); ch.outstats(); o.outstats()
    for i in range(3):
        l_o = o.step(); l_g = ch.step()
        assert np.allclose(l_g[1:9], l_o[1:9], rtol=1e-9, atol=1e-11), (i, l_g, l_o)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-11
    F = ch.download_F()
    assert relerr(F, o.F) < 1e-12 and np.abs(F).max() > 0
    # back to a separable mask on the same handle (coriolis) and off again
    ch.config_coriolis(0.02, 9999999.0, 1.0)
    ch.step()
    assert np.isfinite(ch.download_V().view(np.float64)).all()
    ch.close()
