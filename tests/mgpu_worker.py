"""Worker for the multi-GPU parity test: launched by torchrun with N ranks (one per GPU).
Each rank owns an x-slab (mpi_transpose.f90:214-215), runs the CFL pre-pass and a few RK3 steps
through the C ABI with NCCL all-to-all transposes, and checks its slab against the CPU oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from channel_b200 import Channel, DnsIn, _lib  # noqa: E402
from channel_b200.fields import perturbed_laminar  # noqa: E402
from oracle.channel_oracle import DnsIn as ODnsIn, Oracle, coriolis_force  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    def new_nccl_id():          # one unique id per communicator (= per handle), broadcast from rank 0
        buf = C.create_string_buffer(128)
        if rank == 0:
            _lib.check(lib.chb_get_nccl_unique_id(buf), "id")
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())
    # (grid, couette, library options): direct NVLink stores (default), NCCL all-to-all, row-major velocity
    # buffer + two-lane chunk pipeline
    # default (one lane at these transform sizes), two lanes on green-context SM partitions (the default of the headline
    # grid from 4 GPUs on), two lanes on plain streams with small chunks
    grids = [(31, 16, 16, False, {}), (255, 8, 255, False, {}), (255, 8, 255, False, {"CHB_P2P": "0"}),
             (255, 8, 255, False, {"CHB_TWA": "-1", "CHB_LANES": "1"}), (31, 16, 16, True, {}),
             (255, 8, 255, False, {"CHB_LANES": "2"}), (31, 16, 16, True, {"CHB_LANES": "2", "CHB_WORK_GB": "0.0005"}),
             (255, 8, 255, False, {"CHB_LANES": "2", "CHB_GREEN": "0", "CHB_WORK_GB": "0.02"}),
             (31, 16, 16, True, {"CHB_P2P": "0", "CHB_LANES": "1"})]
    worst = 0.0
    for nx, ny, nz, couette, opts in grids:
        for k in ("CHB_P2P", "CHB_TWA", "CHB_LANES", "CHB_GREEN", "CHB_WORK_GB"):
            os.environ.pop(k, None)
        os.environ.update(opts)
        kw = dict(CPI=False, u0=-1.0, uN=1.0) if couette else {}
        p = DnsIn(nx=nx, ny=ny, nz=nz, re=2000.0, deltat=0.0, cflmax=1.0, **kw)
        o = Oracle(ODnsIn(**{k: getattr(p, k) for k in ODnsIn.__dataclass_fields__}))
        V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, eps=2e-2, couette=couette)
        o.V[:] = V0
        ch = Channel(p, rank=rank, nranks=world, nccl_id=new_nccl_id(), device=local, tables=o)
        sl = slice(ch.nx0, ch.nxN + 1)
        if not opts:
            # Fortran-layout upload (staged through the work arena) with the ranks out of step: the early ranks are
            # already storing their zfwd output into the late ranks' receive buffers while those still upload
            import time
            time.sleep(0.25 * rank)
            ch.upload_V_fortran(np.ascontiguousarray(np.transpose(V0[:, :, sl, :], (0, 2, 3, 1))))
        else:
            ch.upload_V(V0[:, :, sl, :])
        if couette:
            o.set_body_force(coriolis_force(0.02, 9999999.0, 1.0)); ch.config_coriolis(0.02, 9999999.0, 1.0)
        ch.cfl_prepass(); o.cfl_prepass()
        lg = ch.outstats(); lo = o.outstats()
        assert np.allclose(lg, lo, rtol=1e-11, atol=1e-13), (rank, lg, lo)
        for i in range(2):
            lo = o.step(); lg = ch.step()
            assert np.allclose(lg[1:9], lo[1:9], rtol=1e-8, atol=1e-10), (rank, i, lg, lo)
            assert np.allclose(lg[[0, 9, 10]], lo[[0, 9, 10]], rtol=1e-10), (rank, i, lg, lo)
        Vg = ch.download_V()
        for c in range(3):
            ref = o.V[c][:, sl, :]
            err = float(np.abs(Vg[c] - ref).max() / np.abs(o.V[c]).max())
            worst = max(worst, err)
            assert err < 1e-11, (rank, (nx, ny, nz), c, err)
        if not opts:      # one Dati.cart.out written by all ranks at their MPI-IO offsets, read back by all
            from channel_b200.dnsdata import read_restart_file
            path = f"/tmp/chb_mgpu_{os.environ.get('MASTER_PORT', '0')}_{nx}.out"
            ch.save_restart_file(path, async_mode=(nx == 31))
            ch.restart_wait()
            dist.barrier()
            t, Vf = read_restart_file(path, p)
            assert t == ch.time and np.array_equal(Vf[:, sl], ch.download_V_fortran()), (rank, "restart file")
            ch.upload_V(np.zeros_like(Vg))
            ch.read_restart_file(path)
            assert np.array_equal(ch.download_V(), Vg), (rank, "restart read")
            dist.barrier()
            if rank == 0:
                os.remove(path)
        ch.close()
    dist.barrier()
    if rank == 0:
        print(f"MGPU_PARITY_OK world={world} worst_rel_err={worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
