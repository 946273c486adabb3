"""Optional kernel variants (environment switches; DESIGN.md 5b lists what each measured on B200 and which are defaults
for which sizes), each against the kernels it replaces.  The file name sorts last on purpose: with `pytest -x` a problem here cannot hide
the results of the validated suite."""
import numpy as np
import pytest

from channel_b200 import RK1_rai
from tests.helpers import make_pair, relerr

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("nx,ny,nz", [(1023, 8, 3), (511, 8, 4)])
def test_xpass_split_variant(nx, ny, nz, monkeypatch):
    """CHB_XPASS_SPLIT=1: the split x-pass (two threads per innermost butterfly position: 12 instead of 6 warps per SM at
    nxd = 1536, two CTAs of 12 warps instead of three of 6 at nxd = 768) must give the products of the default
    kernel to rounding; also checked against numpy on the CPU emulator (tests/test_fft_emul_cpu.py)."""
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CHB_XPASS_SPLIT", flag)
        p, o, ch, V0 = make_pair(nx, ny, nz, eps=5e-2)
        ch.cfl_prepass(); ch.get_step_scalars()
        ch.buildrhs(RK1_rai, True)
        out[flag] = (ch.download_products(), ch.get_step_scalars()["cfl"])
        if flag == "1":
            Pref = o.convolutions(o.V, False)[..., o.izd]
            for k in range(6):
                assert relerr(out[flag][0][k], Pref[k]) < 1e-12, ("product vs oracle", k)
        ch.close()
    for k in range(6):
        assert relerr(out["1"][0][k], out["0"][0][k]) < 1e-13
    assert abs(out["1"][1] - out["0"][1]) <= 1e-13 * out["0"][1]


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("nx,ny,nz", [(1023, 8, 3), (511, 8, 4), (703, 12, 7)])
def test_xpass_persistent_variant(nx, ny, nz, split, monkeypatch):
    """CHB_XPASS_PERSIST=1: persistent x-pass CTAs that walk over the lines of a launch with the next line's inputs
    prefetched into shared memory (cp.async) and the twiddle tables resident there; with one or two threads per innermost
    butterfly position.  Same arithmetic as the one-CTA-per-line kernel: products equal to rounding, and to the oracle."""
    out = {}
    monkeypatch.setenv("CHB_XPASS_SPLIT", split)
    for flag in ("0", "1"):
        monkeypatch.setenv("CHB_XPASS_PERSIST", flag)
        p, o, ch, V0 = make_pair(nx, ny, nz, eps=5e-2)
        ch.cfl_prepass(); ch.get_step_scalars()
        ch.buildrhs(RK1_rai, True)
        out[flag] = (ch.download_products(), ch.get_step_scalars()["cfl"])
        if flag == "1":
            Pref = o.convolutions(o.V, False)[..., o.izd]
            for k in range(6):
                assert relerr(out[flag][0][k], Pref[k]) < 1e-12, ("product vs oracle", k)
        ch.close()
    for k in range(6):
        assert relerr(out["1"][0][k], out["0"][0][k]) < 1e-13
    assert abs(out["1"][1] - out["0"][1]) <= 1e-13 * out["0"][1]


@pytest.mark.parametrize("nx,ny,nz", [(7, 8, 255), (15, 8, 511), (7, 8, 1023)])
def test_zfwd_direct_variant(nx, ny, nz, monkeypatch):
    """CHB_ZF_DIRECT=1: zfwd4 with stage A reading global memory directly instead of through the TMA staging copy
    (two shared-memory passes per point fewer); same products and fields as the default kernel."""
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CHB_ZF_DIRECT", flag)
        p, o, ch, V0 = make_pair(nx, ny, nz, eps=5e-2)
        ch.cfl_prepass(); ch.get_step_scalars()
        ch.buildrhs(RK1_rai, True)
        out[flag] = (ch.download_products(), ch.get_step_scalars()["cfl"])
        if flag == "1":
            Pref = o.convolutions(o.V, False)[..., o.izd]
            for k in range(6):
                assert relerr(out[flag][0][k], Pref[k]) < 1e-12, ("product vs oracle", k)
        ch.close()
    for k in range(6):
        assert relerr(out["1"][0][k], out["0"][0][k]) < 1e-13
    assert abs(out["1"][1] - out["0"][1]) <= 1e-13 * out["0"][1]


@pytest.mark.parametrize("nx,ny,nz,tpl", [(7, 8, 1023, "128"), (15, 8, 511, "128"), (15, 8, 511, "96")])
def test_zpass_threads_per_line_variants(nx, ny, nz, tpl, monkeypatch):
    """CHB_Z_TPL=128|96: the z passes at nzd = 3072 / 1536 with more threads per line (16 instead of 8, 24 instead of
    16 resident warps per SM, no spills); same products as the default kernels."""
    out = {}
    for flag in ("0", tpl):
        monkeypatch.setenv("CHB_Z_TPL", flag)
        monkeypatch.setenv("CHB_ZF_LPC", "2"); monkeypatch.setenv("CHB_ZB_LPC", "2")
        p, o, ch, V0 = make_pair(nx, ny, nz, eps=5e-2)
        ch.cfl_prepass(); ch.get_step_scalars()
        ch.buildrhs(RK1_rai, True)
        out[flag] = (ch.download_products(), ch.get_step_scalars()["cfl"])
        if flag != "0":
            Pref = o.convolutions(o.V, False)[..., o.izd]
            for k in range(6):
                assert relerr(out[flag][0][k], Pref[k]) < 1e-12, ("product vs oracle", k)
        ch.close()
    for k in range(6):
        assert relerr(out[tpl][0][k], out["0"][0][k]) < 1e-13
    assert abs(out[tpl][1] - out["0"][1]) <= 1e-13 * out["0"][1]


def test_other_wavenumbers_stretching_and_cpi_law():
    """A parity case away from the defaults every other test uses: alfa0, beta0, the stretching parameter a and the
    CPI law (type 0); another wall-normal box is covered on the CPU (tests/test_ydir_emul_cpu.py).  (Not an experimental kernel; it lives here because it was
    added when no GPU was left in the round to run it.)"""
    p, o, ch, V0 = make_pair(19, 40, 13, deltat=0.0, cflmax=0.8, re=1800.0, alfa0=0.8, beta0=1.7, a=2.0,
                             CPI=True, CPI_type=0, gamma=0.3)
    ch.cfl_prepass(); o.cfl_prepass()
    assert np.allclose(ch.outstats(), o.outstats(), rtol=1e-10, atol=1e-12)
    for i in range(5):
        lo = o.step(); lg = ch.step()
        assert np.allclose(lg[1:9], lo[1:9], rtol=1e-8, atol=1e-10), (i, lg, lo)
    Vg = ch.download_V()
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-10
    ch.close()


def test_host_side_body_force_hook_matches_device_path():
    """chb_upload_F: a set_body_force hook evaluated on the host (here the coriolis hook in numpy on the downloaded
    Fortran-layout field) gives the same run as the masked linear device path."""
    from channel_b200 import RK2_rai, RK3_rai
    p, o, ch, V0 = make_pair(15, 32, 10, deltat=2e-3, cflmax=0.0, re=1500.0, couette=True, CPI=False, u0=-1.0, uN=1.0)
    p2, o2, ch2, _ = make_pair(15, 32, 10, deltat=2e-3, cflmax=0.0, re=1500.0, couette=True, CPI=False, u0=-1.0, uN=1.0)
    ch.config_coriolis(0.02, 9999999.0, 1.0)
    ymask = (ch.y <= 1.0) | (ch.y >= p.ymax - 1.0)
    Ff = np.zeros((3, ch2.nxB, 2 * p.nz + 1, p.ny + 3), complex)

    def host_hook():                                  # body_forces/coriolis/coriolis.inc:29-41 on the host
        Vf = ch2.download_V_fortran()
        Ff[0][..., ymask] = -0.02 * Vf[1][..., ymask]
        Ff[1][..., ymask] = 0.02 * Vf[0][..., ymask]
        ch2.upload_F_fortran(Ff)

    host_hook()
    for c in (ch, ch2):
        c.cfl_prepass(); c.outstats()
    for _ in range(2):
        for k, RK in enumerate((RK1_rai, RK2_rai, RK3_rai)):
            ch.set_body_force(); host_hook()
            for c in (ch, ch2):
                c.buildrhs(RK, k == 2); c.linsolve(RK[0] / c.deltat)
    assert np.array_equal(ch.download_V(), ch2.download_V())
    assert np.array_equal(ch.download_F(), np.transpose(ch2.download_F_fortran(), (0, 3, 1, 2)))
    ch.close(); ch2.close()


def test_prefetching_solve_sweeps(monkeypatch):
    """CHB_SOLVE_PF=1: S1 / S3 / S4 issue the loads of the next eight rows before processing them; the arithmetic and
    its order are unchanged, so the fields are bit-identical to the default kernels."""
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CHB_SOLVE_PF", flag)
        p, o, ch, V0 = make_pair(31, 49, 21, deltat=0.0, cflmax=1.0, re=3000.0)     # ny - 1 = 48 rows, ny + 3 = 52 planes
        ch.cfl_prepass(); ch.outstats()
        lines = [ch.step() for _ in range(3)]
        out[flag] = (ch.download_V(), np.array(lines))
        ch.close()
    assert np.array_equal(out["1"][0], out["0"][0])
    assert np.array_equal(out["1"][1], out["0"][1])


def test_convection_velocity_diagnostic(tmp_path):
    """The reference's #ifdef convvel branch of convolutions (dnsdata.f90:515-531,546-549,858-860): uconv and its count
    against the oracle's restatement, and Convvel.cart.<n>.out (save_convvel_file :792-816, outstats :908-913).  cu
    divides by |ust|^2, so entries whose velocity is at rounding level (0/0 in the reference too) are left out."""
    p, o, ch, V0 = make_pair(15, 24, 10, deltat=0.0, cflmax=1.0, re=2500.0, eps=5e-2)
    o.enable_convvel(); ch.enable_convvel()
    ch.cfl_prepass(); o.cfl_prepass()
    ch.outstats(); o.outstats()
    for _ in range(3):
        o.step(); ch.step()
    ug, cnt = ch.get_convvel()
    assert cnt == o.convvel_cnt == 2                          # the first sweep only stores Voldz
    # (where the velocity vanishes - the wall planes - the reference divides 0 by 0: NaN or rounding noise, not compared)
    amp = np.abs(o.Voldz)
    ok = (amp > 1e-5 * amp.max()) & np.isfinite(o.uconv)
    ok[:, :, 0, :] = False                                    # ix = 0 is never accumulated
    assert ok.sum() > 1000
    assert np.abs(ug[ok] - o.uconv[ok]).max() <= 1e-7 * np.abs(o.uconv[ok]).max()
    assert np.all(ug[:, :, 0, :] == 0)
    path = tmp_path / "Convvel.cart.1.out"
    ch.save_convvel_file(path)
    raw = np.fromfile(path, dtype=np.float64)
    assert raw.size == ug.size and np.array_equal(raw.reshape(ug.shape), ug / cnt, equal_nan=True)
    u2, c2 = ch.get_convvel()
    assert c2 == 0 and np.all(u2 == 0)                        # uconv=0; convvel_cnt=0 after the file
    o.step(); ch.step()
    assert ch.get_convvel()[1] == 1
    Vg = ch.download_V()                                      # the diagnostic does not touch the solution
    for c in range(3):
        assert relerr(Vg[c], o.V[c]) < 1e-10
    ch.close()


def test_body_force_reconfiguration_clears_the_old_mask():
    """A hook only assigns F inside its mask and the reference's F starts from zero (dnsdata.f90:146): when the force is
    reconfigured with another mask (or after a host-side F from chb_upload_F), nothing of the previous configuration may
    survive outside the new mask."""
    p, o, ch, V0 = make_pair(15, 24, 10, CPI=False, u0=-1.0, uN=1.0, couette=True)
    ch.config_coriolis(0.02, 9999999.0, 1.0)                   # mask: every z mode, y <= 1 or y >= 1 (all rows)
    F1 = ch.download_F()
    assert np.abs(F1).max() > 0
    ch.config_coriolis(0.02, 2.5, 0.3)                         # narrower: |iz| <= 2, y <= 0.3 or y >= 1.7
    F2 = ch.download_F()
    y = ch.y
    iz = np.arange(-p.nz, p.nz + 1)
    inside = ((y <= 0.3) | (y >= 1.7))[:, None, None] & (np.abs(iz) <= 2)[None, None, :]
    inside = np.broadcast_to(inside, F2.shape[1:])
    assert np.abs(F2[:, ~inside]).max() == 0.0                 # no leftovers of the wide mask
    assert np.abs(F2[:, inside]).max() > 0
    assert np.array_equal(F2[:, inside], F1[:, inside])        # same force where both masks are on
    ch.close()
