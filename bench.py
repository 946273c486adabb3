#!/usr/bin/env python
"""bench.py - RK3 timesteps/s of the channel hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W                  # this repository (CUDA, C ABI)
    python bench.py --impl reference --gpus 1 --steps K --warmup W # reference algorithm on the host cores
    torchrun ... bench.py --gpus N ...                             # N ranks, NCCL all-to-all transposes

A "step" is one full RK3 timestep (3 x [set_body_force,] buildrhs, linsolve; channel.f90:118-167)
on a synthetic perturbed-laminar field.  The workload is a BASELINE.json config:
by default the largest one that fits a single B200 (config 3, nx,ny,nz = 511,512,511) at every N
(strong scaling); --workload selects another (1..5 or nx,ny,nz).

On several GPUs two more legs run outside the timed region of the main one:
  parity_check     : two small grids stepped on all N ranks and compared with the CPU oracle (tests/mgpu_worker.py's
                     check, made visible to the driver) - worst relative error over fields and ranks
  headline_config4 : BASELINE.json's headline grid 1023x1024x1023 (configs[3]) where it fits the N GPUs (N >= 2):
                     ms/step, steps/s, ns/DoF/step, fraction of the max(HBM, NVLink) roofline, kernels, NVLink GB/s

Printed JSON (one line, rank 0): the driver contract plus
  roofline     : dominant kernel, algorithmic HBM bytes / CUDA-event time vs MEASURED_PEAKS.json
  step_roofline: whole step, B_step = 3*M*(ny+1)*(464+288r) bytes / t_step
  kernels      : per-kernel share of the step, from CUDA events on the launching stream
  cpu_baseline : the reference algorithm (oracle/, C + OpenMP) timed on this box's host cores
  e2e          : the same metric through the C ABI with host buffers (upload V, K steps each with
                 its Runtimedata scalars read back, download V), host<->device copies included
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {                     # BASELINE.json configs[0..4]
    "1": dict(nx=16, ny=64, nz=16),
    "2": dict(nx=191, ny=384, nz=189),
    "3": dict(nx=511, ny=512, nz=511),
    "4": dict(nx=1023, ny=1024, nz=1023),
    "5": dict(nx=383, ny=512, nz=383, couette=True),
}
DEFAULT_WORKLOAD = "3"
C16 = 16  # bytes per complex128


def parse_workload(w: str):
    if w in CONFIGS:
        d = dict(CONFIGS[w]); d["name"] = f"config{w}"
        return d
    nx, ny, nz = (int(x) for x in w.split(","))
    return dict(nx=nx, ny=ny, nz=nz, name="custom")


def padded(nx, nz):
    """nxd, nzd: 3/2 rule rounded up to 2^k or 3*2^k (fftFIT, ffts.f90:78-86; dnsdata.f90:123)"""
    def fit(n):
        while True:
            m = n
            while m % 2 == 0:
                m //= 2
            if m in (1, 3):
                return n
            n += 1
    return fit(3 * (nx + 1) // 2), fit(3 * nz)


def workload_text(w, nx, ny, nz, nxd, nzd, couette):
    """config.workload: the same string from both arms of the bench (B200 and reference)"""
    return (f"{w['name']}: turbulent-channel grid nx,ny,nz={nx},{ny},{nz} (nxd,nzd={nxd},{nzd}), perturbed laminar "
            f"{'Couette+coriolis' if couette else 'Poiseuille, CPI'} field, FP64, cflmax=1")


def algorithmic_bytes(nx, ny, nz, nxd, nzd):
    """compulsory HBM bytes per kernel family per RK substep and per step."""
    M = (nx + 1) * (2 * nz + 1)
    r = nzd / (2 * nz + 1)
    npl = ny + 3
    per = {
        "zfwd": (3 + 3 * r) * C16 * M * npl,
        "xpass": (3 * r + 6 * r) * C16 * M * npl,
        "zbwd": (6 * r + 6) * C16 * M * npl,
        "rhs": (11 + 4) * C16 * M * (ny - 1),
        "solve": (2 + 3) * C16 * M * (ny - 1),
    }
    step = 3.0 * M * (ny + 1) * (464 + 288 * r)
    return per, step


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_sm = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for b, n in names.items():
                    if bits & b:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def measured_traffic(kernel, nx, ny, nz, world):
    """DRAM bytes per y-plane of one launch family from the committed `ncu --set full` capture
    (profiles/*_dram_traffic_*.json: dram__bytes_read.sum + dram__bytes_write.sum / planes), for the
    workload and GPU count it was captured on; None otherwise."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_dram_traffic_*.json")), reverse=True):
        try:
            with open(path) as f:
                d = json.load(f)
            if d.get("workload") == f"{nx},{ny},{nz}" and int(d.get("n_gpus", 1)) == world and kernel in d["kernels"]:
                return float(d["kernels"][kernel]["dram_bytes_per_plane"]), os.path.basename(path)
        except Exception:
            continue
    return None, None


FP64_PEAK_TFLOPS = 33.8   # measured FP64 FMA ceiling of one B200 (profiles/README.md, round 1 stage a)
# what each kernel family is limited by according to its ncu capture (profiles/)
LIMITERS = {"xpass": "fp64 pipe + shared-memory crossbar + barriers (not HBM: 0.34 of the HBM roofline at 55 % of the FP64 pipe)",
            "zfwd": "hbm (latency-exposed, 24 % occupancy)", "zbwd": "hbm (latency-exposed, 24 % occupancy)",
            "rhs": "hbm (streaming, 5.2 TB/s of DRAM traffic)", "solve": "hbm (streaming; 15 C moved where 5 C are compulsory)"}
NVLINK_PEAK_GBS = 900.0   # NVLink 5 per direction and GPU (north_star; B200_PROFILING.md)


def nvlink_bytes_per_gpu_step(nx, ny, nz, nzd, world):
    """bytes one GPU sends (= receives) over NVLink per RK3 step: 3 substeps x
    (ny+3) planes x its M/P modes x 9 r C (3 velocity components forward, 6 products back) x the
    (P-1)/P share of every pencil that lives on a peer.  Returns (total, zTOx part, xTOz part)."""
    M = (nx + 1) * (2 * nz + 1)
    r = nzd / (2 * nz + 1)
    per_comp = 3.0 * (M / world) * (ny + 3) * r * C16 * (world - 1) / world
    return 9.0 * per_comp, 3.0 * per_comp, 6.0 * per_comp


def nvlink_report(nx, ny, nz, nzd, world, steps, ms_per_step, kern, direct):
    """Achieved NVLink GB/s per direction and GPU against 900 GB/s.  Direct mode (default): the pack side
    of zTOx / xTOz IS the store loop of zfwd / xpass into peer HBM (CUDA IPC), so the transfer time of a
    transpose is the duration of that kernel (CUDA events); NCCL mode (CHB_P2P=0): the grouped
    ncclSend/ncclRecv all-to-all ("alltoall" timer, both directions together)."""
    total, fwd, bwd = nvlink_bytes_per_gpu_step(nx, ny, nz, nzd, world)
    rep = {"peak_gbs_per_direction": NVLINK_PEAK_GBS, "bytes_per_gpu_step": total,
           "mode": "direct peer stores fused into zfwd/xpass" if direct else "NCCL grouped send/recv all-to-all",
           "step": {"gbs": total / (ms_per_step * 1e-3) / 1e9,
                    "frac": total / (ms_per_step * 1e-3) / 1e9 / NVLINK_PEAK_GBS,
                    "what": "bytes sent per GPU and step / whole step time (overlap-inclusive)"}}
    names = ({"zTOx": (["zfwd"], fwd), "xTOz": (["xpass"], bwd)} if direct
             else {"zTOx+xTOz": (["alltoall"], total)})
    for k, (ns, b) in names.items():
        tms = sum(kern[n][0] for n in ns if n in kern) / steps
        if tms > 0:
            rep[k] = {"carrier": "+".join(ns), "ms_per_step": tms, "gbs": b / (tms * 1e-3) / 1e9,
                      "frac": b / (tms * 1e-3) / 1e9 / NVLINK_PEAK_GBS}
    if "p2p_barrier" in kern:
        rep["barrier_ms_per_step"] = kern["p2p_barrier"][0] / steps
    return rep


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def build_report(args, w, nx, ny, nz, nxd, nzd, world, couette, ms_per_step, kern, launches, clocks, e2e_s, vbytes,
                 dev_bytes, finite, last_line, snap):
    """The JSON line of the B200 arm from the measurements (pure: no GPU needed, unit-tested on the CPU).
    kern: {timer name: (total ms over the timed steps, launches)} from chb_timing_report."""
    dof = 3 * (2 * nx + 1) * (2 * nz + 1) * ny          # README.md:26 convention
    peak, peak_src = measured_peaks()
    per, step_bytes = algorithmic_bytes(nx, ny, nz, nxd, nzd)
    fam = {"zfwd": ["zfwd"], "xpass": ["xpass"], "zbwd": ["zbwd"], "rhs": ["rhs"],
           "solve": ["solve_s1", "solve_s2", "solve_s3", "solve_s4", "solve"]}
    total_kernel_ms = sum(v[0] for v in kern.values()) or 1.0
    kernels = {}
    for f, names in fam.items():
        tms = sum(kern[n][0] for n in names if n in kern)
        if tms <= 0:
            continue
        bytes_total = per[f] / world * 3 * args.steps     # per rank, 3 substeps per step
        kernels[f] = {"ms_per_step": tms / args.steps, "share": tms / total_kernel_ms,
                      "gbs": bytes_total / (tms * 1e-3) / 1e9,
                      "frac": bytes_total / (tms * 1e-3) / 1e9 / peak}
    for n in kern:
        if n.startswith("solve_"):
            kernels["solve"].setdefault("parts_ms_per_step", {})[n] = kern[n][0] / args.steps
        if not any(n in names for names in fam.values()):
            kernels[n] = {"ms_per_step": kern[n][0] / args.steps, "share": kern[n][0] / total_kernel_ms}
    dom = max((k for k in kernels if "gbs" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    nl = sum(kern[n][1] for n in fam[dom] if n in kern)
    tms = sum(kern[n][0] for n in fam[dom] if n in kern)
    bytes_per_launch = per[dom] / world * 3 * args.steps / nl
    ach = bytes_per_launch / (tms / nl * 1e-3) / 1e9
    tpp, tsrc = measured_traffic(dom, nx, ny, nz, world)
    traffic = tpp * (ny + 3) * 3 * args.steps / nl if (tpp and dom in ("zfwd", "xpass", "zbwd")) else None
    # the x-pass is bound by the FP64 pipe and the shared-memory crossbar, not by HBM: report its
    # nominal FFT flops (5 M log2 M per complex transform of the half-length M = nxd, 3 c2r + 6 r2c per z-line,
    # plus ~10 flops per point and transform for the split / merge passes) against the measured FP64 ceiling
    fp64 = None
    if "xpass" in kernels:
        flops_line = 9 * (5.0 * nxd * np.log2(nxd) + 10.0 * nxd) + 12.0 * 2 * nxd
        lines_step = 3.0 * (nzd / world) * (ny + 3)
        tf = float(flops_line * lines_step / (kernels["xpass"]["ms_per_step"] * 1e-3) / 1e12)
        fp64 = {"kernel": "xpass", "nominal_tflops": tf, "peak_tflops": FP64_PEAK_TFLOPS, "frac": tf / FP64_PEAK_TFLOPS,
                "peak_source": "chb_measure_device_peaks on B200, profiles/README.md (33.8 TFLOP/s FMA)"}
    out = {
        "metric": "rk3_timesteps_per_s", "value": 1000.0 / ms_per_step, "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "ns_per_dof_step": ms_per_step * 1e6 / dof,
        "config": {"workload": workload_text(w, nx, ny, nz, nxd, nzd, couette),
                   "dof": dof, "decomposition": f"x-pencils over {world} GPU(s), npy=1",
                   "l2": "state (%.1f GB/GPU) far larger than the 126 MB L2; no flush needed" % (dev_bytes / 1e9),
                   "device_bytes_per_gpu": dev_bytes,
                   # library switches set in the environment; empty = every default
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("CHB_") and k != "CHB_WORKLOAD"}},
        # bound: the contract's roofline is the HBM one (algorithmic bytes / time against the measured copy bandwidth);
        # limiter: what ncu shows the kernel actually waits for
        "roofline": {"bound": "hbm", "limiter": LIMITERS.get(dom, "hbm"), "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                     "bytes_per_launch": bytes_per_launch, "launches": nl, "fp64": fp64},
        "step_roofline": {"bytes_per_step": step_bytes, "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                          "peak": peak * world, "unit": "GB/s",
                          "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world)},
        "kernels": kernels,
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": args.steps / e2e_s, "unit": "steps/s",
                "h2d_bytes_per_step": vbytes * world / args.steps,
                "d2h_bytes_per_step": vbytes * world / args.steps + 8 * 40,
                "what": "chb_upload_V (pinned Fortran-layout V) + K x (chb_buildrhs/chb_linsolve x3 + "
                        "chb_get_step_scalars) + chb_download_V, wall clock"},
        "finite": finite,
        # Runtimedata line of the last timed step (dnsdata.f90:878): time, dudy at both walls (u, w), flow
        # rate x, meanpx, flow rate z, meanpz, cfl*deltat, deltat -- laminar values 3, 3, 0, 0, 2 expected
        "runtimedata_last": [float(v) for v in last_line],
    }
    if world > 1:
        # with P > 1 the step is bounded by max(HBM time, NVLink time), both per GPU
        nv_total, _, _ = nvlink_bytes_per_gpu_step(nx, ny, nz, nzd, world)
        t_hbm = step_bytes / world / (peak * 1e9) * 1e3
        t_nvl = nv_total / (NVLINK_PEAK_GBS * 1e9) * 1e3
        out["step_roofline"].update({"hbm_ms": t_hbm, "nvlink_ms": t_nvl, "roofline_ms": max(t_hbm, t_nvl),
                                     "frac_of_max_hbm_nvlink": max(t_hbm, t_nvl) / ms_per_step})
    if snap:
        out["snapshot"] = snap
    if world > 1:
        out["nvlink"] = nvlink_report(nx, ny, nz, nzd, world, args.steps, ms_per_step, kern,
                                      direct=os.environ.get("CHB_P2P", "1") != "0")
    return out


# ---------------------------------------------------------------------------------------------
def _nccl_id_factory(lib, _lib, rank, world):
    """one NCCL unique id per communicator (= per chb_create), broadcast from rank 0 over torch.distributed"""
    import ctypes as C
    import torch
    import torch.distributed as dist

    def new_id():
        if world == 1:
            return None
        buf = C.create_string_buffer(128)
        if rank == 0:
            _lib.check(lib.chb_get_nccl_unique_id(buf), "chb_get_nccl_unique_id")
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())
    return new_id


def parity_check(rank, world, local_rank, new_id):
    """Multi-rank parity made visible to the driver (the check of tests/mgpu_worker.py): two small grids are stepped on
    all `world` ranks through the C ABI - pencil transposes, barriers and the chunk pipeline included - and every
    rank compares its x-slab with the CPU oracle (test infrastructure, used here as the checker only, outside any timed
    region).  Norm-wise relative error, the way utilities/compare_fields.py:17-54 compares fields."""
    import torch
    import torch.distributed as dist
    from channel_b200 import Channel, DnsIn
    from channel_b200.fields import perturbed_laminar
    from oracle.channel_oracle import DnsIn as ODnsIn, Oracle
    worst, cases = 0.0, []
    t0 = time.perf_counter()
    for nx, ny, nz, steps in ((31, 16, 16, 2), (255, 8, 255, 2)):
        p = DnsIn(nx=nx, ny=ny, nz=nz, re=2000.0, deltat=0.0, cflmax=1.0)
        o = Oracle(ODnsIn(**{k: getattr(p, k) for k in ODnsIn.__dataclass_fields__}))
        V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, eps=2e-2)
        o.V[:] = V0
        ch = Channel(p, rank=rank, nranks=world, nccl_id=new_id(), device=local_rank, tables=o)
        sl = slice(ch.nx0, ch.nxN + 1)
        ch.upload_V(V0[:, :, sl, :])
        ch.cfl_prepass(); o.cfl_prepass()
        lg = ch.outstats(); lo = o.outstats()
        err_lines = float(np.max(np.abs(lg - lo) / np.maximum(np.abs(lo), 1e-300) * (np.abs(lo) > 1e-12)))
        for _ in range(steps):
            lo = o.step(); lg = ch.step()
            m = np.abs(lo[1:9]) > 1e-9
            err_lines = max(err_lines, float(np.max(np.abs(lg[1:9] - lo[1:9])[m] / np.abs(lo[1:9])[m])) if m.any() else 0.0)
        Vg = ch.download_V()
        err = max(float(np.abs(Vg[c] - o.V[c][:, sl, :]).max() / np.abs(o.V[c]).max()) for c in range(3))
        ch.close()
        t = torch.tensor([err, err_lines], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cases.append({"grid": [nx, ny, nz], "steps": steps, "field_rel_err": float(t[0].item()),
                      "runtimedata_rel_err": float(t[1].item())})
        worst = max(worst, float(t[0].item()))
    return {"world": world, "worst_rel_err": worst, "tolerance": 1e-11, "ok": bool(worst < 1e-11), "cases": cases,
            "oracle": "oracle/channel_oracle.py (numpy restatement of the reference; parity unpinned, DESIGN.md)",
            "seconds": time.perf_counter() - t0}


def timed_leg(args, w, rank, world, local_rank, new_id, steps, warmup, pinned, e2e, clock_sampler=True):
    """One workload through the C ABI: upload the synthetic field, CFL pre-pass, `warmup` untimed steps, then `steps`
    steps timed on the device (CUDA events on the launching stream, max over ranks); optionally the end-to-end leg
    with host buffers.  Returns everything build_report needs."""
    import torch
    import torch.distributed as dist
    from channel_b200 import Channel, DnsIn
    from channel_b200.fields import perturbed_laminar_slab

    couette = bool(w.get("couette"))
    p = DnsIn(nx=w["nx"], ny=w["ny"], nz=w["nz"], deltat=0.0, cflmax=1.0,
              CPI=not couette, u0=-1.0 if couette else 0.0, uN=1.0 if couette else 0.0)
    ch = Channel(p, rank=rank, nranks=world, nccl_id=new_id(), device=local_rank)
    if couette:
        ch.config_coriolis(0.02, 9999999.0, 1.0)      # body_forces/coriolis/coriolis.in as shipped
    nx, ny, nz, nxd, nzd = ch.nx, ch.ny, ch.nz, ch.nxd, ch.nzd

    # synthetic perturbed-laminar field of this rank's x-slab, in the Fortran (Dati.cart.out)
    # layout V(iy,iz,ix,c): this is what the driver's read_restart_file would hold
    shape = (3, ch.nxB, 2 * nz + 1, ny + 3)
    if pinned:
        Vf = torch.empty(shape, dtype=torch.complex128).pin_memory()
        Vn = Vf.numpy()
    else:
        Vn = np.empty(shape, dtype=np.complex128)
    perturbed_laminar_slab(Vn, nx, ny, nz, p.alfa0, p.beta0, ch.nx0, ch.nxB, p.a, p.ymin, p.ymax,
                           eps=1e-3, couette=couette)
    vbytes = Vn.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ch.upload_V_fortran(Vn)
    ch.cfl_prepass()
    ch.outstats()                                      # deltat = cflmax / cfl  (channel.f90:116)
    for _ in range(warmup):
        ch.step()

    # ---- timed region: K steps, device time on the launching stream, max over ranks ----------
    ch.timing_enable(True)
    l0 = ch.launch_count()
    sampler = ClockSampler(local_rank) if clock_sampler else None
    barrier()
    if sampler:
        sampler.start()
    ch.stopwatch_begin()
    last_line = None
    for _ in range(steps):
        last_line = ch.step()
    ms = ch.stopwatch_end()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ch.launch_count() - l0
    kern = ch.timing_report()
    ch.timing_enable(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = dict(ch=ch, Vn=Vn, vbytes=vbytes, ms_per_step=ms / steps, kern=kern, launches=launches, clocks=clocks,
               last_line=last_line, couette=couette, dims=(nx, ny, nz, nxd, nzd), barrier=barrier, e2e_s=None)

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    if e2e:
        barrier()
        t0 = time.perf_counter()
        ch.upload_V_fortran(Vn)
        for _ in range(steps):
            ch.step()                                  # includes chb_get_step_scalars D2H per step
        ch.download_V_fortran(out=Vn)
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        out["e2e_s"] = e2e_s
    return out


def headline_leg(args, rank, world, local_rank, new_id):
    """BASELINE.json configs[3], the grid the metric is quoted on: 1023 x 1024 x 1023 on the N GPUs of this run, where
    it fits (5.5 complex per point resident: 189 GB on one GPU - does not fit -, 95 GB per GPU on two)."""
    import torch
    w = parse_workload("4")
    nx, ny, nz = w["nx"], w["ny"], w["nz"]
    nxd, nzd = padded(nx, nz)
    M = (nx + 1) * (2 * nz + 1)
    need = 5.5 * C16 * (ny + 3) * M / world + 11e9          # resident fields + work arena + tables
    free_b, total_b = torch.cuda.mem_get_info()
    host_need = 3.0 * C16 * (ny + 3) * M / world * min(world, 8)     # the ranks of this node build their slabs at once
    try:
        with open("/proc/meminfo") as f:
            host_avail = next(int(l.split()[1]) * 1024 for l in f if l.startswith("MemAvailable"))
    except Exception:
        host_avail = None
    fits = not (need > free_b or (host_avail is not None and host_need > 0.8 * host_avail))
    if world > 1:      # every rank takes the same decision (a rank that skipped alone would leave the others in a collective)
        import torch.distributed as dist
        t = torch.tensor([1 if fits else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        fits = bool(t.item())
    if not fits:
        return {"ran": False, "why": f"needs {need / 1e9:.0f} GB per GPU ({free_b / 1e9:.0f} GB free) and "
                                     f"{host_need / 1e9:.0f} GB of host memory for the synthetic field "
                                     f"({(host_avail or 0) / 1e9:.0f} GB available)"}
    r = timed_leg(args, w, rank, world, local_rank, new_id, args.headline_steps, args.headline_warmup, pinned=False,
                  e2e=False, clock_sampler=True)
    ch = r["ch"]
    hargs = argparse.Namespace(steps=args.headline_steps, warmup=args.headline_warmup)
    rep = build_report(hargs, w, nx, ny, nz, nxd, nzd, world, False, r["ms_per_step"], r["kern"], r["launches"], r["clocks"],
                       1.0, r["vbytes"], ch.device_bytes(), bool(np.isfinite(r["last_line"]).all()), r["last_line"], None)
    ch.close()
    keep = {k: rep[k] for k in ("ms_per_step", "ns_per_dof_step", "step_roofline", "kernels", "nvlink", "gpu_launches",
                                "clocks", "finite", "runtimedata_last", "roofline") if k in rep}
    keep.update({"ran": True, "workload": rep["config"]["workload"], "steps_per_s": rep["value"], "steps": args.headline_steps,
                 "warmup": args.headline_warmup, "n_gpus": world, "dof": rep["config"]["dof"],
                 "device_bytes_per_gpu": rep["config"]["device_bytes_per_gpu"],
                 "target": "north_star: >= 0.60 of the max(HBM, NVLink) roofline on 8 B200 (step_roofline.frac_of_max_hbm_nvlink)"})
    return keep


def run_b200(args):
    import torch
    import torch.distributed as dist
    from channel_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this implementation has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    lib = _lib.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    new_id = _nccl_id_factory(lib, _lib, rank, world)

    # ---- multi-rank parity against the oracle, before anything is timed ---------------------------
    parity = None
    if world > 1 and args.parity_check:
        try:
            parity = parity_check(rank, world, local_rank, new_id)
        except Exception as e:      # a failed check must show up in the line, not kill the measurement
            parity = {"world": world, "ok": False, "error": repr(e)[:300]}

    w = parse_workload(args.workload)
    r = timed_leg(args, w, rank, world, local_rank, new_id, args.steps, args.warmup, pinned=True, e2e=True)
    ch, Vn, barrier = r["ch"], r["Vn"], r["barrier"]
    nx, ny, nz, nxd, nzd = r["dims"]
    ms_per_step, kern, launches, clocks, e2e_s, vbytes = (r[k] for k in ("ms_per_step", "kern", "launches", "clocks", "e2e_s", "vbytes"))
    last_line, couette = r["last_line"], r["couette"]
    finite = bool(np.isfinite(last_line).all())

    # ---- optional: snapshot files from the device-resident field ------------
    snap = None
    if args.snapshot and world == 1:
        d = args.snapshot_dir or ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
        path = os.path.join(d, f"chb_bench_{os.getpid()}.out")
        try:
            barrier()
            t0 = time.perf_counter(); ch.save_restart_file(path, async_mode=False); t_block = time.perf_counter() - t0
            st_b = ch.restart_stats()
            ch.stopwatch_begin(); ch.step(stats=False); ch.step(stats=False); ms_2 = ch.stopwatch_end()
            barrier()
            t0 = time.perf_counter(); ch.save_restart_file(path, async_mode=True); t_call = time.perf_counter() - t0
            ch.stopwatch_begin(); ch.step(stats=False); ch.step(stats=False); ms_2a = ch.stopwatch_end()
            ch.restart_wait(); t_async = time.perf_counter() - t0
            st_a = ch.restart_stats()
            snap = {"bytes": st_b["bytes"], "dir": d,
                    "blocking": {"seconds": t_block, "gbs": st_b["bytes"] / t_block / 1e9,
                                 "device_transposition_ms": st_b["snapshot_ms"]},
                    "async": {"call_returns_after_s": t_call, "file_complete_after_s": t_async,
                              "device_transposition_ms": st_a["snapshot_ms"],
                              "two_steps_ms_alone": ms_2, "two_steps_ms_while_draining": ms_2a}}
        finally:
            if os.path.exists(path):
                os.remove(path)

    out = None
    if rank == 0:
        out = build_report(args, w, nx, ny, nz, nxd, nzd, world, couette, ms_per_step, kern, launches, clocks, e2e_s, vbytes,
                           ch.device_bytes(), finite, last_line, snap)
        if parity is not None:
            out["parity_check"] = parity
    ch.close()
    del Vn, r

    # ---- the headline grid of BASELINE.json on the GPUs of this run (several GPUs only) -------------
    if world > 1 and args.headline and args.workload != "4":
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            head = headline_leg(args, rank, world, local_rank, new_id)
        except Exception as e:
            head = {"ran": False, "why": repr(e)[:300]}
        if rank == 0:
            out["headline_config4"] = head

    if rank == 0:
        if args.cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(w, sample_s=args.cpu_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def cpu_baseline(w, sample_s=20.0, threads=None):
    """The reference algorithm (oracle/channel_oracle_c.c, C + OpenMP, plane-by-plane like
    dnsdata.f90) on this box's host cores (all of the affinity mask, whatever OMP_NUM_THREADS says), on a bounded
    sample of the same grid."""
    from oracle import c_oracle
    return c_oracle.timed_sample(w["nx"], w["ny"], w["nz"], seconds=sample_s, threads=threads)


def run_reference(args):
    """The reference arm: the reference's algorithm on the host cores of this box (oracle/channel_oracle_c.c; the Fortran
    / FFTW / MPI binary cannot be built in this image).  A "step" of this arm is one bounded sample of an RK3
    step of the same workload (the task's definition; with the driver's 25 steps one complete RK substep, a third of a
    step), scaled to a full step: `value` and `ms_per_step` are the extrapolated full-workload figures (`extrapolated: true`, the executed share in `sampled_fraction_of_step`, the
    wall time actually spent per sample in `sample_wall_ms`).  `extrapolation_check` times one COMPLETE, unsampled RK3
    step on config 2 next to the estimate the same sampling gives for it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    w = parse_workload(args.workload)
    nxd, nzd = padded(w["nx"], w["nz"])
    threads = c_oracle.host_cores()
    vals, walls = [], []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(w, sample_s=args.cpu_seconds, threads=threads)
        if i >= args.warmup:
            vals.append(cb["value"])
            walls.append(cb["sample_wall_s"])
    v = float(np.mean(vals))
    cb["value"] = v
    check = None
    if args.extrapolation_check:
        try:
            c2 = parse_workload(args.extrapolation_grid)
            check = c_oracle.timed_full_step(c2["nx"], c2["ny"], c2["nz"], threads=threads)
            check["what"] = (f"one complete unsampled RK3 step of {c2['name']} (co_step, after one warm-up step) against the "
                             "estimate the bounded sampling gives for the same grid")
        except Exception as e:
            check = {"error": repr(e)[:200]}
    out = {"impl": "reference", "metric": "rk3_timesteps_per_s", "value": v, "unit": "steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": workload_text(w, w["nx"], w["ny"], w["nz"], nxd, nzd, bool(w.get("couette"))),
                      "dof": 3 * (2 * w["nx"] + 1) * (2 * w["nz"] + 1) * w["ny"]},
           "extrapolated": True, "sampled_fraction_of_step": cb["sampled_fraction_of_step"],
           "sample_wall_ms": 1000.0 * float(np.mean(walls)),
           "extrapolation_check": check,
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CHB_WORKLOAD", DEFAULT_WORKLOAD))
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--snapshot", action="store_true", help="also time chb_save_restart_file (blocking and asynchronous)")
    ap.add_argument("--snapshot-dir", default=None)
    ap.add_argument("--no-parity-check", dest="parity_check", action="store_false", help="several GPUs: skip the oracle comparison")
    ap.add_argument("--no-headline", dest="headline", action="store_false", help="several GPUs: skip the config-4 leg")
    ap.add_argument("--headline-steps", type=int, default=5)
    ap.add_argument("--headline-warmup", type=int, default=3)
    ap.add_argument("--no-extrapolation-check", dest="extrapolation_check", action="store_false")
    ap.add_argument("--extrapolation-grid", default="2", help="reference arm: grid of the complete-step check (config number or nx,ny,nz)")
    args = ap.parse_args()
    if args.impl == "reference":
        # each "step" of the reference arm is one bounded sample, at best one complete RK substep of the workload (a third of
        # a step: 11 s for config 3 on 16 cores); the whole run stays within about eight minutes
        args.cpu_seconds = min(max(args.cpu_seconds, 20.0), max(3.0, 480.0 / max(1, args.steps + args.warmup)))
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
