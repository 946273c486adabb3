"""channel_b200_run, the C++ restatement of PROGRAM channel (channel.f90:16-193) on top of the C ABI:
a run from dns.in alone must write the same Runtimedata and Dati.cart.out as the Python mirror of the
driver loop (same library underneath), snapshots at the dt_field cadence of outstats
(dnsdata.f90:895-918), and continue an existing Runtimedata from the restart time (get_record)."""
import os
import subprocess

import numpy as np
import pytest

from channel_b200 import Channel, _lib, read_dnsin
from channel_b200.dnsdata import read_restart_file

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(_lib.LIB_PATH), "channel_b200_run")

DNS_IN = """16 64 16            ! nx, ny, nz
0.5 1.0             ! alfa0 beta0
1500                ! ni
1.5 0.0 2.0         ! a, ymin, ymax
.FALSE. 1 0.161436  ! CPI, CPItype, gamma
0.002 0.0           ! meanpx, meanpz
0.0 0.0             ! meanflowx meanflowz
{walls}             ! u0 uN
0.05 0.0 0.0        ! deltat, cflmax, time
0.12 {dt_save} 1000 .TRUE.   ! dt_field, dt_save, t_max, time_from_restart
{nstep}             ! nstep
1                   ! npy
"""


def _rtd(path):
    return np.array([[float(x) for x in l.split()] for l in open(path) if l.strip()])


def test_driver_matches_python_loop_and_restarts(tmp_path):
    d = tmp_path
    (d / "dns.in").write_text(DNS_IN.format(walls="0.0 0.0", dt_save="-1", nstep=6))
    out = subprocess.run([EXE, "--dir", str(d)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Generating initial field" in out.stdout
    rtd = _rtd(d / "Runtimedata")
    assert rtd.shape == (7, 11)                                     # outstats before the loop + 6 steps
    # the same run through the Python mirror of the driver
    p = read_dnsin(str(d / "dns.in"))
    ch = Channel(p)
    V = np.zeros(ch.field_shape(), complex)
    V[0, :, 0, p.nz] = 3 * 0.5 * ch.y * (2 - ch.y)                  # dnsdata.f90:713
    ch.upload_V(V)
    ch.cfl_prepass()
    ch.deltat = p.deltat                                            # cflmax = 0: channel.f90:70-72
    lines = [ch.outstats()] + [ch.step() for _ in range(6)]
    assert np.allclose(rtd, np.array(lines), rtol=1e-14, atol=1e-300)
    t, Vf = read_restart_file(d / "Dati.cart.out", p)
    assert abs(t - ch.time) < 1e-15 and np.array_equal(Vf, ch.download_V_fortran())
    # dt_field = 0.12, deltat = 0.05: snapshots when a multiple of 0.12 falls inside a step (outstats :895-918)
    snaps = sorted(f for f in os.listdir(d) if f.startswith("Dati.cart.") and f != "Dati.cart.out")
    assert snaps == ["Dati.cart.1.out", "Dati.cart.2.out"]
    assert os.path.getsize(d / snaps[0]) == os.path.getsize(d / "Dati.cart.out")
    t1, _ = read_restart_file(d / "Dati.cart.1.out", p)
    assert abs(t1 - 0.10) < 1e-12 or abs(t1 - 0.15) < 1e-12
    ch.close()
    # second run: restarts from Dati.cart.out, finds its time in Runtimedata and continues the file there
    (d / "dns.in").write_text(DNS_IN.format(walls="0.0 0.0", dt_save="0.1", nstep=3))
    out = subprocess.run([EXE, "--dir", str(d)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Reading from file Dati.cart.out" in out.stdout and "starting from time" in out.stdout
    rtd2 = _rtd(d / "Runtimedata")
    assert rtd2.shape == (6 + 1 + 3, 11)                            # the matching record is overwritten
    assert np.array_equal(rtd2[:6], rtd[:6]) and abs(rtd2[6, 0] - rtd[6, 0]) < 1e-15
    assert np.all(np.diff(rtd2[:, 0]) > 0)
    assert "Writing Dati.cart.out at time" in out.stdout            # dt_save cadence (:882-886)
    t2, _ = read_restart_file(d / "Dati.cart.out", p)
    assert abs(t2 - rtd2[-1, 0]) < 1e-15


def test_driver_couette_coriolis(tmp_path):
    """u0=-1, uN=1 walls + coriolis.in as shipped (BASELINE config 5's physics on a small grid)."""
    d = tmp_path
    (d / "dns.in").write_text(DNS_IN.format(walls="-1.0 1.0", dt_save="-1", nstep=2).replace("0.002 0.0", "0.0 0.0  "))
    (d / "coriolis.in").write_text("0.02    ! 2*Ro\n9999999 ! kz_cutoff\n1.0     ! y_threshold\n")
    out = subprocess.run([EXE, "--dir", str(d), "--coriolis"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Coriolis force" in out.stdout
    rtd = _rtd(d / "Runtimedata")
    assert rtd.shape == (3, 11) and np.isfinite(rtd).all()
    p = read_dnsin(str(d / "dns.in"))
    t, Vf = read_restart_file(d / "Dati.cart.out", p)
    assert np.isfinite(Vf.view(np.float64)).all()
    assert abs(Vf[0, 0, p.nz, 1] + 1.0) < 1e-12 and abs(Vf[0, 0, p.nz, p.ny + 1] - 1.0) < 1e-12     # wall velocities of the mean mode
    assert os.path.exists(d / "Dati.cart.1.out") == os.path.exists(d / "Force.cart.1.out")
