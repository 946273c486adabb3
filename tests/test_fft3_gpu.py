"""The register-resident three-stage FFT kernels (zpass3_kernels.cu / xpass3_kernels.cu) serve the
large transform sizes only (nzd in {768,1536,3072}, nxd in {384,768,1536}); these grids are thin
in the other directions so that the oracle stays fast while the new kernels are the ones that run.
Same tolerances as test_parity_gpu.py (1e-12 relative, single step)."""
import os

import numpy as np
import pytest

from channel_b200 import RK1_rai, RK2_rai, RK3_rai
from tests.helpers import make_pair, relerr

pytestmark = pytest.mark.gpu

# (nx, ny, nz) -> (nxd, nzd)
GRIDS = [
    (7, 8, 255),      # nzd = 768
    (15, 8, 511),     # nzd = 1536
    (7, 8, 1023),     # nzd = 3072
    (255, 8, 5),      # nxd = 384
    (511, 8, 4),      # nxd = 768
    (1023, 8, 3),     # nxd = 1536
    (255, 10, 255),   # both large
]


# kernel variants: lines per CTA of zfwd / zbwd (CHB_ZF_LPC, CHB_ZB_LPC) and the tile width of the
# products work buffer (CHB_TW); the first is the default
VARIANTS = [("", "", ""), ("8", "4", "2"), ("2", "8", "0")]


@pytest.mark.parametrize("zf,zb,tw", VARIANTS)
@pytest.mark.parametrize("nx,ny,nz", GRIDS)
def test_fft3_products_and_step(nx, ny, nz, zf, zb, tw, monkeypatch):
    for k, v in (("CHB_ZF_LPC", zf), ("CHB_ZB_LPC", zb), ("CHB_TW", tw)):
        if v:
            monkeypatch.setenv(k, v)
        else:
            monkeypatch.delenv(k, raising=False)
    p, o, ch, V0 = make_pair(nx, ny, nz, eps=5e-2)
    ch.cfl_prepass(); o.cfl_prepass()
    s = ch.get_step_scalars()
    assert abs(s["cfl"] - o.cfl) <= 1e-12 * o.cfl
    o.cfl = 0.0
    for RK, last in ((RK1_rai, False), (RK2_rai, False), (RK3_rai, True)):
        Pref = o.convolutions(o.V, False)[..., o.izd]
        ch.buildrhs(RK, last)
        Pgpu = ch.download_products()
        for k in range(6):
            assert relerr(Pgpu[k], Pref[k]) < 1e-12, ("product", k, relerr(Pgpu[k], Pref[k]))
        o.buildrhs(RK, last)
        lam = RK[0] / p.deltat
        o.linsolve(lam); ch.linsolve(lam)
        Vg = ch.download_V()
        for c in range(3):
            assert relerr(Vg[c], o.V[c]) < 1e-12, ("field", c)
    s = ch.get_step_scalars()
    assert abs(s["cfl"] - o.cfl) <= 1e-12 * o.cfl
    ch.close()


def test_fft3_matches_generic_kernels(monkeypatch):
    """Same inputs through the generic shared-memory passes (CHB_FFT3=0) and the specialised ones."""
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CHB_FFT3", flag)
        p, o, ch, V0 = make_pair(255, 8, 255, eps=5e-2)
        ch.cfl_prepass(); ch.get_step_scalars()
        ch.buildrhs(RK1_rai, True)
        out[flag] = (ch.download_products(), ch.get_step_scalars()["cfl"])
        ch.close()
    for k in range(6):
        assert relerr(out["1"][0][k], out["0"][0][k]) < 1e-13
    assert abs(out["1"][1] - out["0"][1]) <= 1e-13 * out["0"][1]
