"""Snapshot / restart files from the device-resident field (restart_io.cu; save_restart_file
dnsdata.f90:821-848, read_restart_file dnsdata.f90:677-704).  Bit-exact: the file a GPU run writes must
be byte-identical to the reference layout of the same field (here produced by the plain Python
writer from a host copy), and reading it back must reproduce the device field exactly."""
import numpy as np
import pytest

from channel_b200 import Channel, DnsIn, _lib
from channel_b200.dnsdata import read_restart_file, save_restart_file
from channel_b200.fields import perturbed_laminar

pytestmark = pytest.mark.gpu


def _channel(nx, ny, nz, **kw):
    p = DnsIn(nx=nx, ny=ny, nz=nz, re=2000.0, deltat=2e-3, cflmax=0.0, **kw)
    ch = Channel(p)
    V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, eps=1e-2)
    ch.upload_V(V0)
    return p, ch, V0


@pytest.mark.parametrize("nx,ny,nz,chunk_mb", [(16, 64, 16, None), (31, 48, 21, "0.25"), (7, 16, 5, "0.01")])
@pytest.mark.parametrize("async_mode", [False, True])
def test_snapshot_file_is_byte_identical(nx, ny, nz, chunk_mb, async_mode, tmp_path, monkeypatch):
    if chunk_mb:
        monkeypatch.setenv("CHB_IO_CHUNK_MB", chunk_mb)   # many chunks, ragged last chunk
    p, ch, V0 = _channel(nx, ny, nz)
    ch.time = 3.5
    ref = tmp_path / "ref.out"; got = tmp_path / "Dati.cart.out"
    save_restart_file(ref, p, 3.5, ch.download_V_fortran())
    ch.save_restart_file(got, async_mode=async_mode)
    if async_mode:
        ch.restart_wait()
    assert got.read_bytes() == ref.read_bytes()
    st = ch.restart_stats()
    assert st["bytes"] == 3 * (nx + 1) * (2 * nz + 1) * (ny + 3) * 16 and st["total_s"] > 0
    ch.close()


@pytest.mark.parametrize("async_mode", [False, True])
def test_small_work_arena_stages_in_x_slabs(async_mode, tmp_path, monkeypatch):
    """The work arena of the pencil transposes is the staging area of the blocking snapshot, of the restart read and of
    the Fortran-layout transfers; when it is smaller than the field they go through it slab by slab of x-modes."""
    monkeypatch.setenv("CHB_WORK_GB", "0.0005")           # one plane per chunk: room for 12 of the 32 x-modes
    monkeypatch.setenv("CHB_IO_CHUNK_MB", "0.05")
    p, ch, V0 = _channel(31, 48, 21)
    Vf = np.ascontiguousarray(np.transpose(V0, (0, 2, 3, 1)))       # [c][ix][iz][iy]
    assert np.array_equal(ch.download_V_fortran(), Vf)
    ch.upload_V(np.zeros_like(V0)); ch.upload_V_fortran(Vf)
    assert np.array_equal(ch.download_V(), V0)
    ref = tmp_path / "ref.out"; got = tmp_path / "Dati.cart.out"
    ch.time = 1.25
    save_restart_file(ref, p, 1.25, Vf)
    ch.save_restart_file(got, async_mode=async_mode)
    ch.restart_wait()
    assert got.read_bytes() == ref.read_bytes()
    ch.upload_V(np.zeros_like(V0))
    assert ch.read_restart_file(got) == 1.25
    assert np.array_equal(ch.download_V(), V0)
    ch.close()


def test_async_snapshot_is_taken_at_call_time(tmp_path):
    """The time loop continues while the snapshot drains: the file holds the field of the call."""
    p, ch, V0 = _channel(31, 48, 21)
    ch.cfl_prepass(); ch.outstats()
    ch.step()
    Vf = ch.download_V_fortran().copy(); t = ch.time
    ch.save_restart_file(tmp_path / "a.out", async_mode=True)
    for _ in range(2):
        ch.step()                                   # overwrites V, P and everything else on the device
    ch.restart_wait()
    t2, V2 = read_restart_file(tmp_path / "a.out", p)
    assert t2 == t and np.array_equal(V2, Vf)
    assert not np.array_equal(ch.download_V_fortran(), Vf)
    # a second snapshot reuses the buffers
    ch.save_restart_file(tmp_path / "b.out", async_mode=True)
    ch.restart_wait()
    assert np.array_equal(read_restart_file(tmp_path / "b.out", p)[1], ch.download_V_fortran())
    ch.close()


def test_read_restart_file_roundtrip_and_header_check(tmp_path, monkeypatch):
    monkeypatch.setenv("CHB_IO_CHUNK_MB", "0.125")
    p, ch, V0 = _channel(16, 64, 16)
    path = tmp_path / "Dati.cart.out"
    Vf = np.ascontiguousarray(np.transpose(V0, (0, 2, 3, 1)))       # [c][ix][iz][iy]
    save_restart_file(path, p, 12.75, Vf)
    ch.upload_V(np.zeros_like(V0))
    assert ch.read_restart_file(path) == 12.75 and ch.time == 12.75
    assert np.array_equal(ch.download_V(), V0)
    # same step from the restarted field as from the uploaded one
    ch.cfl_prepass(); ch.outstats(); l1 = ch.step()
    ch2 = Channel(p); ch2.upload_V(V0); ch2.time = 12.75
    ch2.cfl_prepass(); ch2.outstats(); l2 = ch2.step()
    assert np.array_equal(l1, l2) and np.array_equal(ch.download_V(), ch2.download_V())
    # metadata mismatch stops with the reference's message (dnsdata.f90:696-703); missing file = code 4
    other = Channel(DnsIn(nx=16, ny=64, nz=16, re=2001.0))
    with pytest.raises(_lib.ChannelB200Error, match="mismatch in metadata"):
        other.read_restart_file(path)
    with pytest.raises(_lib.ChannelB200Error, match="code 4"):
        other.read_restart_file(tmp_path / "nope.out")
    (tmp_path / "short.out").write_bytes(path.read_bytes()[:1000])
    with pytest.raises(_lib.ChannelB200Error):
        ch.read_restart_file(tmp_path / "short.out")
    for c in (ch, ch2, other):
        c.close()


def test_force_snapshot(tmp_path):
    """Force.cart.<n>.out (dnsdata.f90:905-906) from the device-resident body force."""
    p, ch, V0 = _channel(15, 32, 10, CPI=False, u0=-1.0, uN=1.0)
    with pytest.raises(_lib.ChannelB200Error):
        ch.save_restart_file(tmp_path / "f.out", field="F")         # body force not enabled
    ch.config_coriolis(0.02, 9999999.0, 1.0)
    ch.save_restart_file(tmp_path / "f.out", field="F")
    F = ch.download_F()
    t, Ff = read_restart_file(tmp_path / "f.out", p)
    assert np.array_equal(np.transpose(Ff, (0, 3, 1, 2)), F)
    ch.close()
