"""world_size-2 gloo test (CPU) of the multi-GPU host logic: the x-pencil decomposition
(mpi_transpose.f90:214-215), per-rank slab generation, and the block layout of the pencil
transposes zTOx / xTOz (mpi_transpose.f90:50-117) as an all-to-all of contiguous per-peer
blocks - exactly the exchange transpose.cu issues with ncclSend/ncclRecv on the GPUs."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from channel_b200 import _lib


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, nx, nz, nzd, ncomp, nplanes, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = _lib.load()
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert lib.chb_host_decomposition(nx, nzd, world, rank, C.byref(a), C.byref(b), C.byref(c), C.byref(d)) == 0
        nx0, nxN, nz0, nzN = a.value, b.value, c.value, d.value
        nxB, nzB = nxN - nx0 + 1, nzN - nz0 + 1
        idx = lambda *args: lib.chb_host_transpose_index(*args)
        # z side: rank owns x-modes nx0..nxN and all nzd physical z-lines; value encodes (c,pl,izd,ix)
        code = lambda cc, pl, izd, ix: complex(((cc * 100 + pl) * 10000 + izd) * 10000 + ix, rank * 0)
        n = world * ncomp * nplanes * nzB * nxB
        send = np.zeros(n, complex)
        for cc in range(ncomp):
            for pl in range(nplanes):
                for izd in range(nzd):
                    peer, izl = divmod(izd, nzB)
                    for ixl in range(nxB):
                        send[idx(peer, ncomp, cc, nplanes, pl, nzB, izl, nxB, ixl)] = code(cc, pl, izd, nx0 + ixl)
        recv = np.zeros(n, complex)
        ts = torch.from_numpy(send.view(np.float64)); tr = torch.from_numpy(recv.view(np.float64))
        dist.all_to_all_single(tr, ts)                      # the zTOx block exchange
        # x side: rank owns z-lines nz0..nzN and all x-modes 0..nx, read the way xpass_kernel does
        for cc in range(ncomp):
            for pl in range(nplanes):
                for izl in range(nzB):
                    for ix in range(nx + 1):
                        peer, ixl = divmod(ix, nxB)
                        got = recv[idx(peer, ncomp, cc, nplanes, pl, nzB, izl, nxB, ixl)]
                        assert got == code(cc, pl, nz0 + izl, ix), (rank, cc, pl, izl, ix, got)
        # xTOz is the same exchange in reverse: sending recv back must restore send
        back = np.zeros(n, complex)
        dist.all_to_all_single(torch.from_numpy(back.view(np.float64)), tr)
        assert np.array_equal(back, send)
        # per-rank slabs reassemble the full synthetic field
        from channel_b200.fields import perturbed_laminar, perturbed_laminar_slab
        ny = 8
        slab = np.empty((3, nxB, 2 * nz + 1, ny + 3), complex)
        perturbed_laminar_slab(slab, nx, ny, nz, 0.5, 1.0, nx0, nxB)
        parts = [torch.zeros_like(torch.from_numpy(slab.view(np.float64))) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(slab.view(np.float64)))
        full = np.concatenate([p_.numpy().view(np.complex128) for p_ in parts], axis=1)
        ref = np.transpose(perturbed_laminar(nx, ny, nz, 0.5, 1.0), (0, 2, 3, 1))
        assert np.abs(full - ref).max() < 1e-18
        # one Dati.cart.out written by all ranks at the offsets of the MPI-IO view (restart_io.cu's host side)
        import struct, tempfile
        path = os.path.join(tempfile.gettempdir(), f"chb_gloo_{port}.out")
        total = lib.chb_host_restart_file_bytes(nx, ny, nz)
        fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
        os.ftruncate(fd, total)
        if rank == 0:
            hdr = (C.c_ubyte * 68)()
            lib.chb_host_restart_header(nx, ny, nz, 0.5, 1.0, 1e-3, 1.5, 0.0, 2.0, 4.5, hdr)
            os.pwrite(fd, bytes(hdr), 0)
        for cc in range(3):
            os.pwrite(fd, np.ascontiguousarray(slab[cc]).tobytes(), lib.chb_host_restart_offset(nx, ny, nz, nx0, cc))
        os.close(fd)
        dist.barrier()
        raw = open(path, "rb").read()
        assert len(raw) == total and struct.unpack("<3i7d", raw[:68])[:3] == (nx, ny, nz)
        assert np.array_equal(np.frombuffer(raw, dtype=np.complex128, offset=68).reshape(full.shape), full)
        dist.barrier()
        if rank == 0:
            os.remove(path)
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


@pytest.mark.parametrize("world", [2, 4])
def test_pencil_transpose_block_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nx, nz = 7, 3          # nx+1 = 8 modes, nzd = 12: both divisible by 2 and 4
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, nz, 12, 3, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
