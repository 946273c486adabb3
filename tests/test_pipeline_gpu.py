"""The chunk pipeline of the nonlinear term (chb_api.cu, convolutions_all): the planes of a sweep are processed in
chunks, the plane loop of buildrhs follows chunk by chunk with carried accumulators, and with two lanes the kernels
that carry the pencil transposes run on their own stream (and, with CHB_GREEN, their own SM partition) one chunk ahead
of the local kernels.  The arithmetic and its order do not depend on the chunking, so every configuration must give
bit-identical fields and Runtimedata lines; the reference's counterpart is the equality of its blocking and
nonblockingXZ variants (mpi_transpose.f90:149-168)."""
import numpy as np
import pytest

from tests.helpers import make_pair, relerr

pytestmark = pytest.mark.gpu


def _run(monkeypatch, env, steps=3, grid=(31, 48, 21), couette=False):
    for k in ("CHB_WORK_GB", "CHB_LANES", "CHB_GREEN"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    kw = dict(CPI=False, u0=-1.0, uN=1.0) if couette else {}
    p, o, ch, V0 = make_pair(*grid, deltat=0.0, cflmax=1.0, re=3000.0, couette=couette, **kw)
    if couette:
        ch.config_coriolis(0.02, 9999999.0, 1.0)
    ch.cfl_prepass(); ch.outstats()
    lines = [ch.step() for _ in range(steps)]
    out = (ch.download_V(), np.array(lines), ch.download_products())
    ch.close()
    return out


@pytest.mark.parametrize("env", [
    {"CHB_WORK_GB": "0.002", "CHB_LANES": "1"},                       # many chunks, one stream
    {"CHB_WORK_GB": "0.002", "CHB_LANES": "2"},                       # two lanes, two plain streams
    {"CHB_WORK_GB": "0.002", "CHB_LANES": "2", "CHB_GREEN": "72"},    # two lanes on disjoint SM partitions (green contexts)
    {"CHB_LANES": "2"},                                               # two lanes at the default budget: four chunks
])
@pytest.mark.parametrize("couette", [False, True])
def test_chunking_is_bit_identical(env, couette, monkeypatch):
    ref = _run(monkeypatch, {"CHB_LANES": "1"}, couette=couette)      # one chunk holds every plane
    got = _run(monkeypatch, env, couette=couette)
    assert np.array_equal(got[0], ref[0])
    assert np.array_equal(got[1], ref[1])
    assert np.array_equal(got[2], ref[2])


def test_two_lanes_against_oracle(monkeypatch):
    """the overlapped pipeline against the CPU oracle (not only against the sequential one)"""
    monkeypatch.setenv("CHB_WORK_GB", "0.004")
    monkeypatch.setenv("CHB_LANES", "2")
    p, o, ch, V0 = make_pair(47, 40, 32, deltat=0.0, cflmax=1.0, re=3000.0)
    ch.cfl_prepass(); o.cfl_prepass()
    assert np.allclose(ch.outstats(), o.outstats(), rtol=1e-12, atol=1e-14)
    for _ in range(2):
        lo = o.step(); lg = ch.step()
        assert np.allclose(lg[1:9], lo[1:9], rtol=1e-9, atol=1e-11)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-11
    ch.close()


def test_rhs_lives_in_V_between_buildrhs_and_linsolve():
    """dnsdata.f90:667-671: buildrhs leaves the RHS of the eta / D2v equations in V(1:ny-1,:,:,1:2)"""
    from channel_b200 import RK1_rai
    p, o, ch, V0 = make_pair(16, 32, 12)
    ch.cfl_prepass(); o.cfl_prepass()
    rhs_ref = o.buildrhs(RK1_rai, False)
    ch.buildrhs(RK1_rai, False)
    Vg = ch.download_V()
    rhs = ch.download_rhs()
    sl = slice(2, p.ny + 1)
    assert np.array_equal(Vg[:2, sl], rhs[:, sl])
    assert relerr(rhs[0, sl], rhs_ref[0, sl]) < 1e-12 and relerr(rhs[1, sl], rhs_ref[1, sl]) < 1e-12
    assert np.array_equal(Vg[2], V0[2])                       # w untouched until linsolve
    ch.close()
