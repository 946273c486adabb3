"""CPU tests of the host side and of the C-ABI library surface (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from channel_b200 import DnsIn, _lib, read_dnsin
from channel_b200.dnsdata import Tables, format_runtimedata, padded_sizes, read_restart_file, save_restart_file
from channel_b200.fields import perturbed_laminar, perturbed_laminar_slab
from oracle.channel_oracle import DnsIn as ODnsIn, Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("channel_b200.h", "channel_b200_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(chb_[A-Za-z0-9_]+)\s*\(", text))
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 40
    for n in decl:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"
    # the ctypes table binds exactly the declared set
    assert set(_lib.SYMBOLS) == decl, set(_lib.SYMBOLS) ^ decl
    assert lib.chb_version() >= 1000


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device chb_create returns an error and a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.chb_create(C.byref(h), 16, 64, 16, 32, 48, 0.5, 1.0, 1e-3, 1.5, 0.0, 2.0, 0, 1, None, 0)
    assert rc != 0 and len(lib.chb_last_error()) > 0
    from channel_b200 import Channel, ChannelB200Error
    with pytest.raises(ChannelB200Error):
        Channel(DnsIn())


def test_argument_validation_precedes_device_use():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.chb_create(C.byref(h), 16, 64, 16, 33, 48, 0.5, 1.0, 1e-3, 1.5, 0.0, 2.0, 0, 1, None, 0) == 2   # odd nxd
    assert lib.chb_create(C.byref(h), 16, 64, 16, 32, 48, 0.5, 1.0, 1e-3, 1.5, 0.0, 2.0, 0, 2, None, 0) == 2   # 2 !| 17
    assert b"README.md:154" in lib.chb_last_error()
    assert lib.chb_create(C.byref(h), 16, 64, 16, 32, 48, 0.5, 1.0, 1e-3, 1.5, 0.0, 2.0, 0, 9, None, 0) == 2   # more than 8 ranks
    assert lib.chb_create(C.byref(h), 16, 4000, 16, 32, 48, 0.5, 1.0, 1e-3, 1.5, 0.0, 2.0, 0, 1, None, 0) == 2  # mean-mode column > shared memory
    assert b"ny too large" in lib.chb_last_error()
    assert lib.chb_download_products(None, None) != 0 and lib.chb_debug_capture_products(None, 1) != 0
    assert lib.chb_buildrhs(None, None, 1.0, 0) != 0
    assert lib.chb_linsolve(None, 1.0) != 0


def test_padded_sizes_and_decomposition():
    assert padded_sizes(191, 189) == (384, 768) and padded_sizes(1023, 1023) == (1536, 3072)
    lib = _lib.load()
    a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    cover_x, cover_z = [], []
    for r in range(8):   # config 4 on 8 GPUs (mpi_transpose.f90:214-215)
        assert lib.chb_host_decomposition(1023, 3072, 8, r, C.byref(a), C.byref(b), C.byref(c), C.byref(d)) == 0
        cover_x += list(range(a.value, b.value + 1)); cover_z += list(range(c.value, d.value + 1))
    assert cover_x == list(range(1024)) and cover_z == list(range(3072))
    assert lib.chb_host_decomposition(16, 48, 2, 0, C.byref(a), C.byref(b), C.byref(c), C.byref(d)) == 2


@pytest.mark.parametrize("ny", [16, 64, 96])
def test_host_tables_match_oracle(ny):
    """host_tables.cpp restates setup_derivatives/setup_boundary_conditions (dnsdata.f90:241-308);
    the numpy oracle restates them independently."""
    t = Tables(ny, 1.5, 0.0, 2.0)
    o = Oracle(ODnsIn(nx=4, ny=ny, nz=4))
    assert np.allclose(t.y, o.y, rtol=0, atol=1e-15)
    for n in ("d0", "d1", "d2", "d4"):
        ref = getattr(o, n)[2:ny + 1]
        assert np.abs(getattr(t, n) - ref).max() <= 1e-12 * np.abs(ref).max(), n
    assert np.abs(t.D0mat - o.D0mat).max() <= 1e-12 * np.abs(o.D0mat).max()
    for n in ("d140", "d14m1", "d240", "d24m1", "d14n", "d14np1", "d24n", "d24np1",
              "v0bc", "v0m1bc", "vnbc", "vnp1bc", "eta0bc", "eta0m1bc", "etanbc", "etanp1bc"):
        ref = getattr(o, n)
        assert np.abs(getattr(t, n) - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), n


DNS_IN = """\
191 384 189        ! nx, ny, nz
0.5d0 1.0d0        ! alfa0, beta0
12431              ! ni (read as Re; inverted by read_dnsin)
1.5d0 0.0d0 2.0d0  ! a, ymin, ymax
.TRUE. 1 0.161436d0 ! CPI, CPI_type, gamma
0.0 0.0            ! meanpx, meanpz
0.0 0.0            ! meanflowx, meanflowz
0.0 0.0            ! u0, uN
0.0 1.0 0.0        ! deltat, cflmax, time
30.0 -1 7000.0 .TRUE. ! dt_field, dt_save, t_max, time_from_restart
999999             ! nstep
1                  ! npy
"""


def test_read_dnsin(tmp_path):
    f = tmp_path / "dns.in"
    f.write_text(DNS_IN.replace("d0", "e0"))
    p = read_dnsin(str(f))
    assert (p.nx, p.ny, p.nz) == (191, 384, 189) and p.CPI and p.CPI_type == 1
    assert p.re == 12431.0 and abs(p.gamma - 0.161436) < 1e-15 and p.cflmax == 1.0 and p.npy == 1
    f.write_text("1 2 3\n")
    with pytest.raises(ValueError):
        read_dnsin(str(f))


def test_restart_file_roundtrip_and_header_check(tmp_path):
    """Dati.cart.out: 3 x int32 + 7 x float64 header, then V(iy,iz,ix,c) (dnsdata.f90:683-695,830-846)."""
    p = DnsIn(nx=5, ny=8, nz=3, re=100.0)
    V = np.random.default_rng(1).standard_normal((3, 6, 7, 11)) + 0j
    path = str(tmp_path / "Dati.cart.out")
    save_restart_file(path, p, 1.25, V)
    assert os.path.getsize(path) == 68 + V.nbytes
    t, V2 = read_restart_file(path, p)
    assert t == 1.25 and np.array_equal(V, V2)
    with pytest.raises(ValueError):
        read_restart_file(path, DnsIn(nx=5, ny=8, nz=3, re=101.0))
    assert len(format_runtimedata(np.arange(11.0)).split()) == 11


def test_slab_generator_matches_full_field():
    nx, ny, nz = 15, 16, 7
    V = perturbed_laminar(nx, ny, nz, 0.5, 1.0, couette=True)
    for nx0, nxB in ((0, 16), (0, 8), (8, 8), (12, 4)):
        out = np.empty((3, nxB, 2 * nz + 1, ny + 3), complex)
        perturbed_laminar_slab(out, nx, ny, nz, 0.5, 1.0, nx0, nxB, couette=True)
        assert np.abs(out - np.transpose(V[:, :, nx0:nx0 + nxB], (0, 2, 3, 1))).max() < 1e-18
    # Hermitian symmetry on the ix=0 line, v(0,0)=0
    assert np.array_equal(V[:, :, 0, :nz], np.conj(V[:, :, 0, :nz:-1])) and np.all(V[1, :, 0, nz] == 0)


def test_bench_algorithmic_bytes_match_survey():
    import bench
    per, step = bench.algorithmic_bytes(1023, 1024, 1023, 1536, 3072)
    assert abs(step - 5.78e12) / 5.78e12 < 0.01             # 5.78 TB/step at config 4
    per, step = bench.algorithmic_bytes(191, 384, 189, 384, 768)
    assert abs(step - 88e9) / 88e9 < 0.01                   # 88 GB/step at config 2
    assert abs(3 * sum(per.values()) - step) / step < 0.02  # per-kernel rows add up to the step total


def test_bench_nvlink_and_hbm_byte_models():
    """bench.py's byte models reproduce SURVEY.md 8(d): config 4 moves 5.78 TB of HBM traffic per step
    and 152.6 GB per GPU over NVLink at P=8; config 2 moves 88 GB per step."""
    import bench
    _, step4 = bench.algorithmic_bytes(1023, 1024, 1023, 1536, 3072)
    assert abs(step4 / 1e12 - 5.78) < 0.01
    _, step2 = bench.algorithmic_bytes(191, 384, 189, 384, 768)
    assert abs(step2 / 1e9 - 88.0) < 0.5
    tot, fwd, bwd = bench.nvlink_bytes_per_gpu_step(1023, 1024, 1023, 3072, 8)
    assert abs(tot / 1e9 - 152.6) < 0.1 and abs(fwd * 2 - bwd) < 1e-6 * bwd
    rep = bench.nvlink_report(511, 512, 511, 1536, 2, 3, 138.0, {"zfwd": (300.0, 15), "xpass": (500.0, 15)}, True)
    assert rep["zTOx"]["carrier"] == "zfwd" and 0 < rep["step"]["frac"] < 1


def test_restart_file_layout_host_helpers(tmp_path):
    """Dati.cart.out layout (dnsdata.f90:683-695,830-846): the C helpers of restart_io.cu agree with the
    Python writer (channel_b200.dnsdata.save_restart_file) and with the way the reference's own
    utilities read the header (utilities/compare_fields.py:17-18: 3 int32, then 7 float64 at offset 12)."""
    import struct
    from channel_b200 import DnsIn
    from channel_b200.dnsdata import save_restart_file, read_restart_file, HEADER_FMT
    lib = _lib.load()
    nx, ny, nz = 5, 8, 3
    p = DnsIn(nx=nx, ny=ny, nz=nz, re=1234.5)
    buf = (C.c_ubyte * 68)()
    assert lib.chb_host_restart_header(nx, ny, nz, p.alfa0, p.beta0, 1.0 / p.re, p.a, p.ymin, p.ymax, 7.25, buf) == 0
    raw = bytes(buf)
    assert raw == struct.pack(HEADER_FMT, nx, ny, nz, p.alfa0, p.beta0, 1.0 / p.re, p.a, p.ymin, p.ymax, 7.25)
    ints = np.frombuffer(raw, dtype=np.int32, count=3); reals = np.frombuffer(raw, dtype=np.float64, count=7, offset=12)
    assert list(ints) == [nx, ny, nz] and reals[-1] == 7.25 and reals[2] == 1.0 / p.re
    # offsets: component c of the slab starting at nx0 inside the C-order [3][nx+1][2nz+1][ny+3] array
    rng = np.random.default_rng(1)
    V = rng.standard_normal((3, nx + 1, 2 * nz + 1, ny + 3)) + 1j * rng.standard_normal((3, nx + 1, 2 * nz + 1, ny + 3))
    path = tmp_path / "Dati.cart.out"
    save_restart_file(path, p, 7.25, V)
    data = path.read_bytes()
    assert len(data) == lib.chb_host_restart_file_bytes(nx, ny, nz)
    for c in range(3):
        for nx0 in (0, 2, 3):
            off = lib.chb_host_restart_offset(nx, ny, nz, nx0, c)
            got = np.frombuffer(data, dtype=np.complex128, count=(2 * nz + 1) * (ny + 3), offset=off)
            assert np.array_equal(got.reshape(2 * nz + 1, ny + 3), V[c, nx0])
    t, V2 = read_restart_file(path, p)
    assert t == 7.25 and np.array_equal(V2, V)
    with pytest.raises(ValueError):
        read_restart_file(path, DnsIn(nx=nx, ny=ny, nz=nz, re=999.0))
    assert lib.chb_save_restart_file(None, b"x", 0.0, 0, 0) != 0 and lib.chb_read_restart_file(None, b"x", None) != 0


SHIPPED_DNS_IN = """191       384    189                     ! nx, ny, nz
0.5       1.0                            ! alfa0 beta0  
12431                                    ! ni
1.5       0.0    2.0                     ! a, ymin, ymax
.TRUE.    1      0.161436                ! CPI, CPItype, gamma
0.0       0.0                            ! meanpx, meanpz
0.0       0.0                            ! meanflowx meanflowz
0.0       0.0                            ! u0 uN
0.0       1.0    0.0                     ! deltat, cflmax, time
30       -1      7000        .TRUE.      ! dt_field, dt_save, t_max, time_from_restart
999999                                  ! nstep
1                                       ! npy (no. processes y)
"""


def test_cpp_driver_reads_dns_in_and_runtimedata(tmp_path):
    """channel_b200_run (C++ restatement of PROGRAM channel) parses the repo's shipped dns.in exactly like
    read_dnsin (dnsdata.f90:98-125) / the Python mirror, and positions an existing Runtimedata like
    get_record (dnsdata.f90:181-218); --check-input needs no GPU."""
    import json, subprocess
    from channel_b200 import read_dnsin
    from channel_b200.dnsdata import padded_sizes, format_runtimedata
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "channel_b200_run")
    assert os.path.exists(exe), "build() makes channel_b200/lib/channel_b200_run"
    (tmp_path / "dns.in").write_text(SHIPPED_DNS_IN)
    out = json.loads(subprocess.check_output([exe, "--dir", str(tmp_path), "--check-input"], text=True))
    p = read_dnsin(str(tmp_path / "dns.in"))
    assert (out["nxd"], out["nzd"]) == padded_sizes(p.nx, p.nz) == (384, 768)
    for k in ("nx", "ny", "nz", "alfa0", "beta0", "a", "ymin", "ymax", "CPI_type", "gamma", "meanpx", "meanpz", "meanflowx",
              "meanflowz", "u0", "uN", "deltat", "cflmax", "time", "dt_field", "dt_save", "t_max", "nstep", "npy"):
        assert out[k] == getattr(p, k), k
    assert out["ni"] == 1.0 / p.re and bool(out["CPI"]) is p.CPI and bool(out["time_from_restart"]) is p.time_from_restart
    assert out["rtd_exists"] == 0
    # Fortran spellings: D exponents, commas, lower-case logicals; restart at time 0.3 inside an existing Runtimedata
    (tmp_path / "dns.in").write_text(SHIPPED_DNS_IN.replace("12431 ", "1.2431D4").replace(".TRUE.    1", "f, 0,")
                                     .replace("0.0       1.0    0.0   ", "1.0d-2, 0.0, 0.3"))
    lines = [format_runtimedata([0.1 * i] + [0.0] * 9 + [0.1]) for i in range(6)]
    (tmp_path / "Runtimedata").write_text("\n".join(lines) + "\n")
    out = json.loads(subprocess.check_output([exe, "--dir", str(tmp_path), "--check-input"], text=True))
    assert out["ni"] == 1.0 / 12431.0 and out["CPI"] == 0 and out["CPI_type"] == 0 and out["deltat"] == 0.01 and out["time"] == 0.3
    assert out["rtd_exists"] == 1 and out["rtd_found"] == 1 and out["rtd_keep_lines"] == 3 and out["rtd_deltat"] == 0.1
    # missing values stop the run like a failed list-directed READ
    (tmp_path / "dns.in").write_text("\n".join(SHIPPED_DNS_IN.splitlines()[:5]))
    r = subprocess.run([exe, "--dir", str(tmp_path), "--check-input"], capture_output=True, text=True)
    assert r.returncode != 0 and "expected 12 data lines" in r.stderr


@pytest.mark.parametrize("world", [1, 2, 8])
def test_bench_report_assembly(world):
    """bench.build_report (the JSON line of the B200 arm) from synthetic measurements: every key of the driver
    contract is there and the numbers are consistent; no GPU involved."""
    import json, types
    import bench
    args = types.SimpleNamespace(steps=4, warmup=3, cpu_baseline=False, cpu_seconds=1.0)
    w = bench.parse_workload("3")
    nx, ny, nz, nxd, nzd = 511, 512, 511, 768, 1536
    ms = 225.8 / world
    kern = {"zfwd": (21.8 * 4 / world, 60), "xpass": (76.1 * 4 / world, 60), "zbwd": (46.5 * 4 / world, 60),
            "rhs": (38.1 * 4 / world, 12), "solve_s1": (12.0 * 4 / world, 24), "solve_s2": (16.3 * 4 / world, 24),
            "solve_s3": (5.9 * 4 / world, 12), "solve_s4": (8.6 * 4 / world, 12), "mean_mode": (13.4 * 4, 12)}
    if world > 1:
        kern["p2p_barrier"] = (0.8, 120)
    out = bench.build_report(args, w, nx, ny, nz, nxd, nzd, world, False, ms, kern, 288 * 4, {"sm_mhz": 1860, "sm_max_mhz": 1965,
                             "reasons": [], "samples": 9}, 1.39, 3 * 16 * 512 // world * 1023 * 515, 58e9 / world, True,
                             np.arange(11.0), None)
    if world > 1:
        out["nvlink"] = bench.nvlink_report(nx, ny, nz, nzd, world, args.steps, ms, kern, direct=True)
    line = json.loads(json.dumps(out))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["n_gpus"] == world and line["dtype"] == "f64" and "workload" in line["config"] and "model" not in line["config"]
    assert abs(line["value"] - 1000.0 / ms) < 1e-9 and line["vs_baseline"] is None
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["kernel"] == "xpass" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.3 < r["frac"] < 0.4 and 0.3 < r["fp64"]["frac"] < 0.45          # the measured round-1 numbers: 0.35 and 0.39
    assert abs(line["step_roofline"]["frac"] - 0.488) < 0.01
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["kernels"]["solve"]["parts_ms_per_step"]["solve_s2"] == pytest.approx(16.3 / world)
    assert line["runtimedata_last"] == list(map(float, range(11)))
    if world > 1:
        assert line["nvlink"]["zTOx"]["carrier"] == "zfwd" and line["nvlink"]["barrier_ms_per_step"] == pytest.approx(0.2)
        sr = line["step_roofline"]
        assert sr["roofline_ms"] == max(sr["hbm_ms"], sr["nvlink_ms"]) and 0 < sr["frac_of_max_hbm_nvlink"] < 1
        if world == 8:
            assert sr["nvlink_ms"] > sr["hbm_ms"]          # NVLink overtakes HBM from P = 4 on


def test_bench_reference_arm_runs_on_the_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): one JSON line with the same
    metric / unit, "impl": "reference", its own cpu_baseline description and a zero-copy e2e."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # launched the way torchrun launches it: OMP_NUM_THREADS=1 in the environment must not reach the CPU arm
    env1 = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "1", "--steps", "1",
                        "--warmup", "0", "--cpu-seconds", "1", "--extrapolation-grid", "31,32,31"], env=env1,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rk3_timesteps_per_s" and d["unit"] == "steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["extrapolated"] is True and 0 < d["sampled_fraction_of_step"] <= 1 and d["sample_wall_ms"] > 0
    chk = d["extrapolation_check"]
    assert chk["grid"] == [31, 32, 31] and chk["full_step_s"] > 0 and chk["finite"] and 0.2 < chk["estimate_over_full"] < 5
    import bench
    assert d["config"]["workload"] == bench.workload_text(bench.parse_workload("1"), 16, 64, 16, 32, 48, False)
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # other ranks of a torchrun launch exit without work and without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _c_prototypes(header_text):
    """{name: [(is_pointer, is_handle)]} of every `int|long long|const char* chb_*(...)` prototype of the header"""
    import re
    text = re.sub(r"/\*.*?\*/", " ", header_text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(chb_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        lst = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                lst.append(("*" in a, a.startswith("chb_handle ") or a.startswith("void* ")))
        protos[name] = lst
    return protos


def test_fortran_shim_matches_the_c_header():
    """fortran/channel_b200_mod.f90 (the iso_c_binding interfaces a maintainer adds to the reference) against
    include/channel_b200.h: every entry point of the Fortran-facing part of the ABI has an interface, every
    BIND(C, name=...) names an exported symbol, the argument counts agree, and an argument is passed by VALUE exactly
    when the C parameter is a scalar or the handle (channel.f90:137-139 needs chb_vetaTOuvw / chb_computeflowrate)."""
    import re
    from channel_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "channel_b200.h")).read()
    protos = _c_prototypes(header)
    fortran_part = _c_prototypes(header.split("test / diagnostics accessors")[0])
    src = open(os.path.join(root, "fortran", "channel_b200_mod.f90")).read()
    src = re.sub(r"&\s*\n\s*", " ", src)                 # join continuation lines
    src = re.sub(r"!.*", "", src)                        # strip comments
    bound = {}
    for m in re.finditer(r"FUNCTION\s+(\w+)\s*\(([^)]*)\)\s*BIND\(C,\s*name=\"(\w+)\"\)(.*?)END FUNCTION", src, flags=re.S):
        fname, args, cname, body = m.group(1), m.group(2), m.group(3), m.group(4)
        assert fname == cname, (fname, cname)
        args = [a.strip().lower() for a in args.split(",") if a.strip()]
        by_value = set()
        for decl in re.finditer(r"^[^\n]*,\s*VALUE\s*::\s*([^\n]*)$", body, flags=re.M):
            by_value |= {re.sub(r"\(.*?\)", "", v).strip().lower() for v in decl.group(1).split(",")}
        bound[cname] = [(a in by_value) for a in args]
    assert len(bound) >= 30
    exported = set(_lib.SYMBOLS)
    for name, vals in bound.items():
        assert name in protos, f"{name}: bound in the Fortran shim but not declared in channel_b200.h"
        assert name in exported, f"{name}: not exported by libchannel_b200.so"
        cargs = protos[name]
        assert len(vals) == len(cargs), (name, len(vals), len(cargs))
        for i, ((is_ptr, is_handle), v) in enumerate(zip(cargs, vals)):
            want_value = (not is_ptr) or is_handle
            assert v == want_value, f"{name} argument {i + 1}: VALUE={v}, C parameter {'scalar/handle' if want_value else 'pointer'}"
    missing = sorted(set(fortran_part) - set(bound) - {"chb_upload_V_planes", "chb_download_V_planes"})   # device-layout transfers: tests only
    assert not missing, f"no Fortran interface for {missing}"
