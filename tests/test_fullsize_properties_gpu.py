"""BASELINE config 2 at its full size (nx,ny,nz = 191,384,189: the repo's shipped dns.in grid) on the GPU,
checked through size-independent properties instead of a full-size oracle run:

* scaling the field by 2 scales the six dealiased products by exactly 4 (every operation of the FFT passes
  is linear and a power-of-two factor is exact in binary floating point): BIT-EXACT equality;
* the products of real physical fields are Hermitian on the ix = 0 line;
* after an RK3 step the field is divergence free in the scheme's own sense, ia*u + ib*w + D_y v = 0 with the
  compact first derivative (COMPLEXderiv, dnsdata.f90:339-373), checked with the oracle's derivative on a
  random sample of columns; the wall nodes carry no-slip; the ix = 0 line stays Hermitian;
* the Runtimedata line is finite and the CPI forcing / flow rate stay at the laminar values within the
  perturbation amplitude.
The same grid exercises the specialised kernels for nxd = 384 and nzd = 768 with many chunks of planes."""
import numpy as np
import pytest

from channel_b200 import Channel, DnsIn, RK1_rai
from channel_b200.fields import perturbed_laminar
from oracle.channel_oracle import DnsIn as ODnsIn, Oracle

pytestmark = pytest.mark.gpu
NX, NY, NZ = 191, 384, 189


def test_config2_full_size_properties(monkeypatch):
    monkeypatch.setenv("CHB_WORK_GB", "1.0")            # several chunks of planes per substep
    p = DnsIn(nx=NX, ny=NY, nz=NZ, deltat=0.0, cflmax=1.0)          # dns.in as shipped: Re, CPI type 1, gamma 0.161436
    V0 = perturbed_laminar(NX, NY, NZ, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, eps=1e-3)
    ch = Channel(p)
    assert ch.nxd == 384 and ch.nzd == 768
    ch.capture_products()

    # ---- products: exact quadratic scaling, Hermitian symmetry on ix = 0 ---------------------------
    ch.upload_V(V0)
    ch.cfl_prepass(); s1 = ch.get_step_scalars()
    ch.buildrhs(RK1_rai, True)
    P1 = ch.download_products()
    cfl1 = ch.get_step_scalars()["cfl"]
    ch.upload_V(2.0 * V0)
    ch.buildrhs(RK1_rai, True)
    P2 = ch.download_products()
    cfl2 = ch.get_step_scalars()["cfl"]
    assert np.isfinite(P1.view(np.float64)).all() and np.abs(P1).max() > 0
    assert np.array_equal(P2, 4.0 * P1)                 # bit-exact
    assert cfl2 == 2.0 * cfl1 and cfl1 > 0              # cfl is linear in the velocity (dnsdata.f90:552-556)
    del P2
    for k in range(6):
        line = P1[k][:, 0, :]                           # [iy, iz+nz] at ix = 0
        err = np.abs(line - np.conj(line[:, ::-1])).max()
        assert err <= 1e-13 * np.abs(P1[k]).max(), ("hermitian", k, err)
    del P1
    ch.capture_products(False)

    # ---- one RK3 step ----------------------------------------------------------------------------
    ch.upload_V(V0)
    ch.cfl_prepass(); ch.outstats()
    line = ch.step()
    assert np.isfinite(line).all()
    assert abs(line[5] - 2.0) < 1e-3                    # flow rate of U = 1.5 y (2-y)
    assert abs(line[1] - 3.0) < 0.1 and abs(line[2] - 3.0) < 0.1   # wall shear dU/dy = 3 (both walls, sign convention of outstats)
    assert 0 < line[9] <= 1.0 + 1e-12                   # cfl*deltat = cflmax
    V = ch.download_V()
    ch.close()
    assert np.isfinite(V.view(np.float64)).all()
    # no-slip at both walls (iy = 0 and ny <-> plane indices 1 and ny+1); the mean mode carries u0 = uN = 0 too
    for c in range(3):
        assert np.abs(V[c][1]).max() < 1e-11 and np.abs(V[c][NY + 1]).max() < 1e-11, ("wall", c)
    # Hermitian symmetry of the ix = 0 line
    for c in range(3):
        l0 = V[c][:, 0, :]
        assert np.abs(l0 - np.conj(l0[:, ::-1])).max() <= 1e-13 * max(1.0, np.abs(l0).max())
    # discrete continuity on a sample of columns
    o = Oracle(ODnsIn(nx=2, ny=NY, nz=2, re=p.re))      # the tables depend on ny, a, ymin, ymax only
    rng = np.random.default_rng(7)
    ix = rng.integers(0, NX + 1, 400); iz = rng.integers(0, 2 * NZ + 1, 400)
    keep = ~((ix == 0) & (iz == NZ))
    ix, iz = ix[keep], iz[keep]
    u, v, w = (np.ascontiguousarray(V[c][:, ix, iz]) for c in range(3))
    vy = o.COMPLEXderiv_full(v.copy())
    ia = 1j * p.alfa0 * ix; ib = 1j * p.beta0 * (iz - NZ)
    div = ia[None] * u + vy + ib[None] * w
    scale = max(np.abs(vy).max(), np.abs(ia[None] * u).max())
    assert np.abs(div).max() <= 1e-10 * scale, (np.abs(div).max(), scale)
