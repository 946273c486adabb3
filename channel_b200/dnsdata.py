"""Host-side mirror of the reference's driver-facing interface for the hot path.

Names follow MODULE dnsdata / PROGRAM channel (dnsdata.f90, channel.f90): read_dnsin,
setup_derivatives (+ setup_boundary_conditions), read/save_restart_file, buildrhs, linsolve,
outstats, and the RK3 time loop.  All numerics run in libchannel_b200.so through the C ABI
(include/channel_b200.h); this module only marshals arguments, exactly what the iso_c_binding
shim (fortran/channel_b200_mod.f90) does for the Fortran driver.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import struct

import numpy as np

from . import _lib

RK1_rai = (120.0 / 32.0, 2.0, 0.0)                  # dnsdata.f90:70
RK2_rai = (120.0 / 8.0, 50.0 / 8.0, 34.0 / 8.0)     # dnsdata.f90:71
RK3_rai = (120.0 / 20.0, 90.0 / 20.0, 50.0 / 20.0)  # dnsdata.f90:72


@dataclasses.dataclass
class DnsIn:
    """The 12 lines of dns.in (dnsdata.f90:110-122).  `re` is line 3; ni = 1/re."""
    nx: int = 16
    ny: int = 64
    nz: int = 16
    alfa0: float = 0.5
    beta0: float = 1.0
    re: float = 12431.0
    a: float = 1.5
    ymin: float = 0.0
    ymax: float = 2.0
    CPI: bool = True
    CPI_type: int = 1
    gamma: float = 0.161436
    meanpx: float = 0.0
    meanpz: float = 0.0
    meanflowx: float = 0.0
    meanflowz: float = 0.0
    u0: float = 0.0
    uN: float = 0.0
    deltat: float = 0.0
    cflmax: float = 1.0
    time: float = 0.0
    dt_field: float = 30.0
    dt_save: float = -1.0
    t_max: float = 7000.0
    time_from_restart: bool = True
    nstep: int = 999999
    npy: int = 1


def _fortran_logical(tok: str) -> bool:
    t = tok.strip().strip(".").upper()
    if t in ("TRUE", "T"):
        return True
    if t in ("FALSE", "F"):
        return False
    raise ValueError(f"bad logical in dns.in: {tok!r}")


def read_dnsin(path: str = "dns.in") -> DnsIn:
    """read_dnsin (dnsdata.f90:98-125): list-directed reads, `!` comments ignored."""
    rows = []
    with open(path) as f:
        for line in f:
            body = line.split("!")[0].replace(",", " ").split()
            if body:
                rows.append(body)
    if len(rows) < 12:
        raise ValueError(f"{path}: expected 12 data lines, found {len(rows)}")
    d = DnsIn()
    d.nx, d.ny, d.nz = (int(x) for x in rows[0][:3])
    d.alfa0, d.beta0 = (float(x) for x in rows[1][:2])
    d.re = float(rows[2][0])
    d.a, d.ymin, d.ymax = (float(x) for x in rows[3][:3])
    d.CPI = _fortran_logical(rows[4][0]); d.CPI_type = int(rows[4][1]); d.gamma = float(rows[4][2])
    d.meanpx, d.meanpz = (float(x) for x in rows[5][:2])
    d.meanflowx, d.meanflowz = (float(x) for x in rows[6][:2])
    d.u0, d.uN = (float(x) for x in rows[7][:2])
    d.deltat, d.cflmax, d.time = (float(x) for x in rows[8][:3])
    d.dt_field, d.dt_save, d.t_max = (float(x) for x in rows[9][:3])
    d.time_from_restart = _fortran_logical(rows[9][3])
    d.nstep = int(rows[10][0])
    d.npy = int(rows[11][0])
    return d


def padded_sizes(nx: int, nz: int):
    lib = _lib.load()
    a, b = C.c_int(), C.c_int()
    lib.chb_host_padded_sizes(nx, nz, C.byref(a), C.byref(b))
    return a.value, b.value


class Tables:
    """setup_derivatives + setup_boundary_conditions (dnsdata.f90:241-308), host side (C++)."""

    def __init__(self, ny, a, ymin, ymax):
        lib = _lib.load()
        self.ny = ny
        self.y = np.zeros(ny + 3)
        self.d0 = np.zeros((ny - 1, 5)); self.d1 = np.zeros((ny - 1, 5))
        self.d2 = np.zeros((ny - 1, 5)); self.d4 = np.zeros((ny - 1, 5))
        self.D0mat = np.zeros((ny + 1, 5))
        t = _lib.HostTables()
        for n in ("y", "d0", "d1", "d2", "d4", "D0mat"):
            setattr(t, n, getattr(self, n).ctypes.data_as(_lib.c_double_p))
        rc = lib.chb_host_setup_tables(ny, a, ymin, ymax, C.byref(t))
        if rc:
            raise _lib.ChannelB200Error(f"chb_host_setup_tables failed ({rc})")
        self.c = t
        for n in ("d140", "d14m1", "d240", "d24m1", "d14n", "d14np1", "d24n", "d24np1",
                  "v0bc", "v0m1bc", "vnbc", "vnp1bc", "eta0bc", "eta0m1bc", "etanbc", "etanp1bc"):
            setattr(self, n, np.array(list(getattr(t, n))))


def _dp(a):
    return a.ctypes.data_as(_lib.c_double_p)


class Channel:
    """One rank of the solver: owns a chb_handle (one GPU)."""

    def __init__(self, p: DnsIn, rank: int = 0, nranks: int = 1, nccl_id: bytes | None = None,
                 device: int = 0, tables=None):
        self.lib = _lib.load()
        self.p = p
        self.nx, self.ny, self.nz = p.nx, p.ny, p.nz
        self.nxd, self.nzd = padded_sizes(p.nx, p.nz)
        self.ni = 1.0 / p.re                                    # dnsdata.f90:115
        self.rank, self.nranks = rank, nranks
        self.h = C.c_void_p()
        _lib.check(self.lib.chb_create(C.byref(self.h), p.nx, p.ny, p.nz, self.nxd, self.nzd,
                                       p.alfa0, p.beta0, self.ni, p.a, p.ymin, p.ymax,
                                       rank, nranks, nccl_id, device), "chb_create")
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.chb_get_decomposition(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.nx0, self.nxN, self.nz0, self.nzN = a.value, b.value, c.value, d.value
        self.nxB = self.nxN - self.nx0 + 1
        # host tables: computed on the host exactly as the Fortran driver would, then passed in
        self.tab = tables if tables is not None else Tables(p.ny, p.a, p.ymin, p.ymax)
        self.set_tables(self.tab)
        self.y = self.tab.y
        self.deltat, self.cflmax, self.time = p.deltat, p.cflmax, p.time
        self.meanpx, self.meanpz = p.meanpx, p.meanpz
        self.corrpx = self.corrpz = 0.0
        self.fr = np.zeros(3)
        self.istep = 0
        _lib.check(self.lib.chb_set_forcing(self.h, p.meanpx, p.meanpz, p.meanflowx, p.meanflowz,
                                            int(p.CPI), p.CPI_type, p.gamma), "chb_set_forcing")
        _lib.check(self.lib.chb_set_wall_velocity(self.h, p.u0, p.uN), "chb_set_wall_velocity")
        self.bodyforce = False

    # -- tables -------------------------------------------------------------------------------
    def set_tables(self, t):
        """t: object with the attributes of chb_set_tables (Tables or the oracle)."""
        ny = self.ny
        def rows(a):   # accept [(ny-1),5] or the oracle's [(ny+3),5]
            a = np.ascontiguousarray(a, dtype=np.float64)
            return np.ascontiguousarray(a[2:ny + 1]) if a.shape[0] == ny + 3 else a
        keep = [np.ascontiguousarray(t.y, dtype=np.float64), rows(t.d0), rows(t.d1), rows(t.d2), rows(t.d4)]
        keep += [np.ascontiguousarray(getattr(t, n), dtype=np.float64) for n in (
            "d140", "d14m1", "d240", "d24m1", "d14n", "d14np1", "d24n", "d24np1",
            "v0bc", "v0m1bc", "vnbc", "vnp1bc", "eta0bc", "eta0m1bc", "etanbc", "etanp1bc", "D0mat")]
        _lib.check(self.lib.chb_set_tables(self.h, *[_dp(a) for a in keep]), "chb_set_tables")

    # -- fields -------------------------------------------------------------------------------
    def field_shape(self):
        return (3, self.ny + 3, self.nxB, 2 * self.nz + 1)

    def upload_V(self, V):
        """V in the device layout [c, iy+1, ix-nx0, iz+nz] (complex128)."""
        V = np.ascontiguousarray(V, dtype=np.complex128)
        assert V.shape == self.field_shape(), (V.shape, self.field_shape())
        _lib.check(self.lib.chb_upload_V_planes(self.h, V.ctypes.data), "chb_upload_V_planes")

    def download_V(self):
        V = np.empty(self.field_shape(), np.complex128)
        _lib.check(self.lib.chb_download_V_planes(self.h, V.ctypes.data), "chb_download_V_planes")
        return V

    def upload_V_fortran(self, Vf):
        """Vf: Fortran-ordered V(-1:ny+1,-nz:nz,nx0:nxN,1:3) i.e. C-order [c][ix][iz][iy]."""
        Vf = np.ascontiguousarray(Vf, dtype=np.complex128)
        assert Vf.shape == (3, self.nxB, 2 * self.nz + 1, self.ny + 3)
        _lib.check(self.lib.chb_upload_V(self.h, Vf.ctypes.data), "chb_upload_V")

    def download_V_fortran(self, out=None):
        Vf = out if out is not None else np.empty((3, self.nxB, 2 * self.nz + 1, self.ny + 3), np.complex128)
        assert Vf.flags.c_contiguous and Vf.dtype == np.complex128
        _lib.check(self.lib.chb_download_V(self.h, Vf.ctypes.data), "chb_download_V")
        return Vf

    def download_rhs(self):
        r = np.empty((2,) + self.field_shape()[1:], np.complex128)
        _lib.check(self.lib.chb_download_rhs(self.h, r.ctypes.data), "chb_download_rhs")
        return r

    def capture_products(self, on: bool = True):
        """diagnostics: keep the spectral products of all planes of the following buildrhs sweeps (they normally
        exist one chunk of planes at a time)"""
        _lib.check(self.lib.chb_debug_capture_products(self.h, int(on)), "chb_debug_capture_products")

    def download_products(self):
        r = np.empty((6,) + self.field_shape()[1:], np.complex128)
        _lib.check(self.lib.chb_download_products(self.h, r.ctypes.data), "chb_download_products")
        return r

    def download_F(self):
        F = np.empty(self.field_shape(), np.complex128)
        _lib.check(self.lib.chb_download_F_planes(self.h, F.ctypes.data), "chb_download_F_planes")
        return F

    # -- body force ---------------------------------------------------------------------------
    def config_body_force_linear(self, A, mask_y, mask_z, exclude_mean=False):
        A = np.ascontiguousarray(A, np.float64).reshape(9)
        my = np.ascontiguousarray(mask_y, np.float64); mz = np.ascontiguousarray(mask_z, np.float64)
        assert my.shape == (self.ny + 3,) and mz.shape == (2 * self.nz + 1,)
        _lib.check(self.lib.chb_set_body_force_linear(self.h, 1, _dp(A), _dp(my), _dp(mz), int(exclude_mean)),
                   "chb_set_body_force_linear")
        self.bodyforce = True
        self.set_body_force()       # config_body_force ends with set_body_force (coriolis.inc:27)

    def config_coriolis(self, omega2, kz_cutoff, y_threshold_bot):
        """body_forces/coriolis/coriolis.inc:4-41 as a masked linear force."""
        y_threshold_top = self.p.ymax - y_threshold_bot
        iz_thr = min(self.nz, int(np.floor(kz_cutoff / self.p.beta0)))
        my = ((self.y <= y_threshold_bot) | (self.y >= y_threshold_top)).astype(np.float64)
        iz = np.arange(-self.nz, self.nz + 1)
        mz = (np.abs(iz) <= iz_thr).astype(np.float64)
        A = np.zeros((3, 3)); A[1, 0] = omega2; A[0, 1] = -omega2
        self.config_body_force_linear(A, my, mz, exclude_mean=False)

    def config_body_force_linear_yz(self, A, mask_yz, exclude_mean=False):
        """general mask(iy, iz), shape (ny+3, 2nz+1)"""
        A = np.ascontiguousarray(A, np.float64).reshape(9)
        m = np.ascontiguousarray(mask_yz, np.float64)
        assert m.shape == (self.ny + 3, 2 * self.nz + 1)
        _lib.check(self.lib.chb_set_body_force_linear_yz(self.h, 1, _dp(A), _dp(m), int(exclude_mean)),
                   "chb_set_body_force_linear_yz")
        self.bodyforce = True
        self.set_body_force()

    def _am_pieces(self, lambdaz_f):
        iz_f = int(np.floor((2.0 * np.pi / lambdaz_f) / (self.p.beta0 / 1000.0) + 0.5))   # NINT, am_f1.inc:7
        yp = np.where(self.y > 1, self.p.ymax - self.y, self.y) * 1000.0              # am_f1.inc:20
        iz = np.arange(-self.nz, self.nz + 1)
        return iz_f, yp, iz

    def config_am_f1(self, lambdaz_f=500.0, amp=1000.0):
        """body_forces/am_f1/am_f1.inc: F = -amp V where lambda_z+ > 2.3 (y+)^2 and |iz| <= iz_f, mean mode excluded."""
        iz_f, yp, iz = self._am_pieces(lambdaz_f)
        with np.errstate(divide="ignore"):
            lzp = np.where(iz == 0, 1e10, 2 * np.pi / (self.p.beta0 * np.abs(iz)) * 1000)
        mask = (lzp[None, :] > 2.3 * yp[:, None] ** 2) & (np.abs(iz) <= iz_f)[None, :]
        self.config_body_force_linear_yz(-amp * np.eye(3), mask.astype(np.float64), exclude_mean=True)

    def config_am_butterfly(self, lambdaz_f=500.0, amp=1000.0):
        """body_forces/am_butterfly/am_butterfly.inc: two boxes, (|iz| <= iz_f, y+ <= 60) and (|iz| > iz_f, y+ > 60)."""
        iz_f, yp, iz = self._am_pieces(lambdaz_f)
        inner = (np.abs(iz) <= iz_f)[None, :]
        mask = (inner & (yp <= 60)[:, None]) | (~inner & (yp > 60)[:, None])
        self.config_body_force_linear_yz(-amp * np.eye(3), mask.astype(np.float64), exclude_mean=True)

    def upload_F_fortran(self, Ff):
        """generic body-force path: F evaluated on the host, Fortran layout [c][ix][iz][iy] (chb_upload_F)"""
        Ff = np.ascontiguousarray(Ff, dtype=np.complex128)
        assert Ff.shape == (3, self.nxB, 2 * self.nz + 1, self.ny + 3)
        _lib.check(self.lib.chb_upload_F(self.h, Ff.ctypes.data), "chb_upload_F")
        self.bodyforce = True

    def download_F_fortran(self):
        Ff = np.empty((3, self.nxB, 2 * self.nz + 1, self.ny + 3), np.complex128)
        _lib.check(self.lib.chb_download_F(self.h, Ff.ctypes.data), "chb_download_F")
        return Ff

    def set_body_force(self):
        _lib.check(self.lib.chb_set_body_force(self.h), "chb_set_body_force")

    # -- hot path -----------------------------------------------------------------------------
    def buildrhs(self, ODE, compute_cfl: bool):
        ode = np.array(ODE, dtype=np.float64)
        _lib.check(self.lib.chb_buildrhs(self.h, _dp(ode), self.deltat, int(compute_cfl)), "chb_buildrhs")

    def linsolve(self, lam: float):
        _lib.check(self.lib.chb_linsolve(self.h, lam), "chb_linsolve")

    def cfl_prepass(self):
        """channel.f90:95-115."""
        if self.deltat == 0:
            self.deltat = 1.0
        _lib.check(self.lib.chb_cfl_prepass(self.h), "chb_cfl_prepass")

    def sync(self):
        _lib.check(self.lib.chb_sync(self.h), "chb_sync")

    def get_step_scalars(self):
        cfl = C.c_double(); cpx = C.c_double(); cpz = C.c_double(); mpx = C.c_double(); mpz = C.c_double()
        fr = np.zeros(3); Ulo = np.zeros(5); Uhi = np.zeros(5); Wlo = np.zeros(5); Whi = np.zeros(5)
        _lib.check(self.lib.chb_get_step_scalars(self.h, C.byref(cfl), _dp(fr), C.byref(cpx), C.byref(cpz),
                                                 C.byref(mpx), C.byref(mpz), _dp(Ulo), _dp(Uhi), _dp(Wlo), _dp(Whi)),
                   "chb_get_step_scalars")
        return dict(cfl=cfl.value, fr=fr, corrpx=cpx.value, corrpz=cpz.value, meanpx=mpx.value, meanpz=mpz.value,
                    U_lo=Ulo, U_hi=Uhi, W_lo=Wlo, W_hi=Whi)

    def outstats(self):
        """dnsdata.f90:853-880: the Runtimedata line (11 columns)."""
        s = self.get_step_scalars()
        runtime_global = s["cfl"]
        if self.cflmax > 0:
            self.deltat = self.cflmax / runtime_global
        self.fr, self.corrpx, self.corrpz = s["fr"], s["corrpx"], s["corrpz"]
        self.meanpx, self.meanpz = s["meanpx"], s["meanpz"]
        t = self.tab
        dudy0 = np.sum(t.d140 * s["U_lo"]); dwdy0 = np.sum(t.d140 * s["W_lo"])
        dudyN = -np.sum(t.d14n * s["U_hi"]); dwdyN = -np.sum(t.d14n * s["W_hi"])
        return np.array([self.time, dudy0, dudyN, dwdy0, dwdyN,
                         self.fr[0] + self.corrpx * self.fr[2], self.meanpx + self.corrpx,
                         self.fr[1] + self.corrpz * self.fr[2], self.meanpz + self.corrpz,
                         runtime_global * self.deltat, self.deltat])

    def step(self, stats: bool = True):
        """One RK3 step of the time loop (channel.f90:118-167)."""
        self.istep += 1
        for k, RK in enumerate((RK1_rai, RK2_rai, RK3_rai)):
            self.time = self.time + 2.0 / RK[0] * self.deltat
            if self.bodyforce:
                self.set_body_force()
            self.buildrhs(RK, k == 2)
            self.linsolve(RK[0] / self.deltat)
        return self.outstats() if stats else None

    def stopwatch_begin(self):
        _lib.check(self.lib.chb_stopwatch_begin(self.h), "chb_stopwatch_begin")

    def stopwatch_end(self) -> float:
        ms = C.c_double()
        _lib.check(self.lib.chb_stopwatch_end(self.h, C.byref(ms)), "chb_stopwatch_end")
        return ms.value

    def rk3_step(self):
        """chb_rk3_step: the three substeps in one C call (no Python in the loop)."""
        _lib.check(self.lib.chb_rk3_step(self.h, self.deltat), "chb_rk3_step")

    def device_bytes(self):
        return int(self.lib.chb_device_bytes(self.h))

    def launch_count(self):
        return int(self.lib.chb_launch_count(self.h))

    def timing_enable(self, on=True):
        _lib.check(self.lib.chb_timing_enable(self.h, int(on)), "chb_timing_enable")

    def timing_report(self):
        cap, stride = 64, 64
        names = C.create_string_buffer(cap * stride)
        ms = np.zeros(cap); n = np.zeros(cap, dtype=np.int64)
        k = self.lib.chb_timing_report(self.h, names, stride, _dp(ms), n.ctypes.data_as(C.POINTER(C.c_longlong)), cap)
        out = {}
        for i in range(min(k, cap)):
            nm = names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
            out[nm] = (float(ms[i]), int(n[i]))
        return out

    # -- restart / snapshot files (restart_io.cu) --------------------------------------------
    def save_restart_file(self, path, field: str = "V", async_mode: bool = False):
        """save_restart_file (dnsdata.f90:821-848) from the device-resident field; collective over ranks."""
        _lib.check(self.lib.chb_save_restart_file(self.h, os.fsencode(path), self.time, {"V": 0, "F": 1}[field],
                                                  int(async_mode)), "chb_save_restart_file")

    def restart_wait(self):
        _lib.check(self.lib.chb_restart_wait(self.h), "chb_restart_wait")

    def restart_stats(self):
        b, ms, s = C.c_double(), C.c_double(), C.c_double()
        _lib.check(self.lib.chb_restart_stats(self.h, C.byref(b), C.byref(ms), C.byref(s)), "chb_restart_stats")
        return dict(bytes=b.value, snapshot_ms=ms.value, total_s=s.value)

    def read_restart_file(self, path):
        """read_restart_file (dnsdata.f90:677-704): loads this rank's slab, sets and returns `time`."""
        t = C.c_double()
        _lib.check(self.lib.chb_read_restart_file(self.h, os.fsencode(path), C.byref(t)), "chb_read_restart_file")
        self.time = t.value
        return self.time

    # -- convection-velocity diagnostic (convvel.cu; #ifdef convvel of the reference) -----------------
    def enable_convvel(self, on: bool = True):
        _lib.check(self.lib.chb_set_convvel(self.h, int(on)), "chb_set_convvel")

    def get_convvel(self):
        """(uconv [3][ny+3][nx+1][nzd] as save_convvel_file orders it, convvel_cnt)"""
        u = np.empty((3, self.ny + 3, self.nx + 1, self.nzd))
        n = C.c_longlong()
        _lib.check(self.lib.chb_get_convvel(self.h, _dp(u), C.byref(n)), "chb_get_convvel")
        return u, n.value

    def save_convvel_file(self, path):
        _lib.check(self.lib.chb_save_convvel_file(self.h, os.fsencode(path)), "chb_save_convvel_file")

    def close(self):
        if self.h:
            self.lib.chb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# -- Dati.cart.out (dnsdata.f90:677-720, 821-848) ---------------------------------------------
HEADER_FMT = "<3i7d"          # nx,ny,nz, alfa0,beta0,ni,a,ymin,ymax,time   (68 bytes)


def save_restart_file(path, p: DnsIn, time: float, V_fortran):
    """V_fortran: C-order [3][nx+1][2nz+1][ny+3] complex128 (= Fortran V(iy,iz,ix,c))."""
    with open(path, "wb") as f:
        f.write(struct.pack(HEADER_FMT, p.nx, p.ny, p.nz, p.alfa0, p.beta0, 1.0 / p.re, p.a, p.ymin, p.ymax, time))
        np.ascontiguousarray(V_fortran, dtype=np.complex128).tofile(f)


def read_restart_file(path, p: DnsIn):
    """Returns (time, V_fortran); aborts on header mismatch like dnsdata.f90:696-703."""
    with open(path, "rb") as f:
        hdr = struct.unpack(HEADER_FMT, f.read(struct.calcsize(HEADER_FMT)))
        nx, ny, nz, alfa0, beta0, ni, a, ymin, ymax, time = hdr
        if (nx, ny, nz) != (p.nx, p.ny, p.nz) or (alfa0, beta0, a, ymin, ymax) != (p.alfa0, p.beta0, p.a, p.ymin, p.ymax) \
                or ni != 1.0 / p.re:
            raise ValueError("ERROR: mismatch in metadata between restart file and dns.in. Stopping.")
        V = np.fromfile(f, dtype=np.complex128, count=3 * (nx + 1) * (2 * nz + 1) * (ny + 3))
    return time, V.reshape(3, nx + 1, 2 * nz + 1, ny + 3)


def format_runtimedata(line) -> str:
    """WRITE(101,*) list-directed line of outstats (dnsdata.f90:878)."""
    return " ".join(f"{v:24.16E}" for v in line)
