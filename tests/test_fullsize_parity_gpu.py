"""Parity at the full size of the BASELINE configurations against the C restatement of the reference
(oracle/channel_oracle_c.c: plane loop of dnsdata.f90 with its ring buffers, per-column banded solves; C99 + OpenMP,
a few seconds per RK3 step of config 2 on the box's host cores).  Tolerances are the north star's: fields after one
step 1e-12 relative, after 10 steps 1e-9, Runtimedata columns 2..9 1e-8; "relative" is norm-wise per field, the way
utilities/compare_fields.py:17-54 compares two Dati.cart.out files.  The oracle is parity-unpinned."""
import types

import numpy as np
import pytest

from channel_b200 import Channel, DnsIn
from channel_b200.fields import perturbed_laminar
from oracle.c_oracle import COracle, host_cores, load as load_c_oracle
from oracle.channel_oracle import DnsIn as ODnsIn

pytestmark = pytest.mark.gpu


def _pair(nx, ny, nz, couette=False, eps=1e-3, **kw):
    load_c_oracle().co_global_threads(host_cores())
    fields = dict(nx=nx, ny=ny, nz=nz, deltat=0.0, cflmax=1.0)
    if couette:
        fields.update(CPI=False, u0=-1.0, uN=1.0)
    fields.update(kw)
    p = DnsIn(**fields)
    o = COracle(ODnsIn(**{k: getattr(p, k) for k in ODnsIn.__dataclass_fields__}))
    V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, eps=eps, couette=couette)
    o.set_V(V0)
    # the oracle's coefficient tables on both sides (the driver passes its own to chb_set_tables)
    names = ("y", "d0", "d1", "d2", "d4", "D0mat", "d140", "d14m1", "d240", "d24m1", "d14n", "d14np1", "d24n", "d24np1",
             "v0bc", "v0m1bc", "vnbc", "vnp1bc", "eta0bc", "eta0m1bc", "etanbc", "etanp1bc")
    tab = types.SimpleNamespace(**{n: o.table(n) for n in names})
    ch = Channel(p, tables=tab)
    ch.upload_V(V0)
    del V0
    if couette:
        o.set_coriolis(0.02, 9999999.0, 1.0)
        ch.config_coriolis(0.02, 9999999.0, 1.0)
    return p, o, ch


def _fields_err(ch, o):
    Vg = ch.download_V()
    Vo = o.get_V()
    return max(float(np.abs(Vg[c] - Vo[c]).max() / np.abs(Vo[c]).max()) for c in range(3))


def _run(p, o, ch, steps):
    o.cfl_prepass(); ch.cfl_prepass()
    lo = o.outstats(); lg = ch.outstats()
    assert np.allclose(lg, lo, rtol=1e-11, atol=1e-13), (lg, lo)
    errs = []
    for i in range(steps):
        lo = o.step(); lg = ch.step()
        assert np.isfinite(lg).all()
        assert np.allclose(lg[1:9], lo[1:9], rtol=1e-8, atol=1e-10), (i, lg, lo)          # Runtimedata cols 2..9
        assert np.allclose(lg[[0, 9, 10]], lo[[0, 9, 10]], rtol=1e-9), (i, lg, lo)        # time, cfl*dt, dt
        if i == 0 or i == steps - 1:
            errs.append(_fields_err(ch, o))
    return errs


def test_config2_full_size_one_and_ten_steps():
    """BASELINE configs[1]: the shipped dns.in grid 191 x 384 x 189 (alfa0 0.5, beta0 1, Re 12431, CPI type 1)."""
    p, o, ch = _pair(191, 384, 189)
    assert (ch.nxd, ch.nzd) == (384, 768)
    e1, e10 = _run(p, o, ch, 10)
    ch.close(); o.close()
    assert e1 < 1e-12, e1
    assert e10 < 1e-9, e10


def test_config5_physics_couette_coriolis():
    """BASELINE configs[4] physics (u0 = -1, uN = 1, body_forces/coriolis) on its x-z grid 383 x 383 with 64 planes:
    nxd = 768, nzd = 1536, the transform sizes of the full configuration."""
    p, o, ch = _pair(383, 64, 383, couette=True)
    assert (ch.nxd, ch.nzd) == (768, 1536)
    e1, e3 = _run(p, o, ch, 3)
    ch.close(); o.close()
    assert e1 < 1e-12, e1
    assert e3 < 1e-10, e3


def test_config3_grid_thin_in_y():
    """BASELINE configs[2] x-z grid 511 x 511 (nxd 768, nzd 1536) with 32 planes: one step at 1e-12."""
    p, o, ch = _pair(511, 32, 511)
    e1, = _run(p, o, ch, 1)[:1]
    ch.close(); o.close()
    assert e1 < 1e-12, e1
