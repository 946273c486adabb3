// restart_io.cu - Dati.cart.out snapshots straight from / into the device-resident field
// (save_restart_file dnsdata.f90:821-848, read_restart_file dnsdata.f90:677-720).
//
// File format (the reference's MPI-IO view, mpi_transpose.f90:243-258): a 68-byte header
//   int32 nx,ny,nz ; float64 alfa0,beta0,ni,a,ymin,ymax,time
// followed by the global array V(-1:ny+1,-nz:nz,0:nx,1:3) in Fortran order, i.e. C order
// [c][ix][iz][iy] complex128.  A rank that owns x-modes nx0..nxN holds, for every component, one
// contiguous block of nxB*(2nz+1)*(ny+3) complex at element offset (c*(nx+1)+nx0)*(2nz+1)*(ny+3):
// the subarray types of the reference collapse to three pwrite()/pread() ranges per rank, so no MPI-IO
// is needed and several ranks (one process per GPU) write the same file concurrently.
//
// Save pipeline: (1) the device transposes its layout [c][iy][ix][iz] into file order (a tiled 2-D
// transpose per component) -- blocking mode into the products buffer P, which is dead between
// time steps; asynchronous mode into a dedicated snapshot buffer, after which the time loop may
// continue at once; (2) a copy stream drains the snapshot in chunks through two pinned host
// buffers; (3) the calling thread (blocking) or a writer thread (asynchronous) pwrite()s chunk k
// while chunk k+1 is in flight over PCIe.  At 1024^3 a snapshot is 103 GB (13 GB per GPU on 8 GPUs):
// the step time hides it completely in asynchronous mode.
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

#include "../../include/channel_b200.h"
#include "../../include/channel_b200_host.h"
#include "chb_internal.h"

#define CHB_RESTART_HEADER_BYTES 68

struct RestartIO {
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t snap_done = nullptr, copied[2] = {nullptr, nullptr};
    char* pinned[2] = {nullptr, nullptr};
    size_t chunk_bytes = 0;
    cplx* snap = nullptr;          // asynchronous mode: device copy of the field in file order
    std::thread worker;
    bool worker_active = false;
    int worker_rc = 0;
    std::string worker_err;
    // measurements of the last save
    double bytes = 0, t_snapshot_ms = 0, t_total_s = 0;
};

// ---- host-only pieces (also exported for the CPU tests) ----------------------------------------
extern "C" int chb_host_restart_header(int nx, int ny, int nz, double alfa0, double beta0, double ni, double a,
                                       double ymin, double ymax, double time, unsigned char* out68) {
    if (!out68) return 2;
    const int ints[3] = {nx, ny, nz};
    const double reals[7] = {alfa0, beta0, ni, a, ymin, ymax, time};
    memcpy(out68, ints, 12);              // [nx,ny,nz], MPI_INTEGER          dnsdata.f90:835
    memcpy(out68 + 12, reals, 56);        // 7 x MPI_DOUBLE_PRECISION         dnsdata.f90:836
    return 0;
}

// byte offset in the file of component c of the x-slab starting at nx0   (writeview_type, mpi_transpose.f90:249-252)
extern "C" long long chb_host_restart_offset(int nx, int ny, int nz, int nx0, int c) {
    const long long col = (long long)(2 * nz + 1) * (ny + 3);
    return CHB_RESTART_HEADER_BYTES + 16LL * ((long long)c * (nx + 1) + nx0) * col;
}

extern "C" long long chb_host_restart_file_bytes(int nx, int ny, int nz) {
    return chb_host_restart_offset(nx, ny, nz, 0, 3);
}

static int write_all(int fd, const char* p, size_t n, off_t off, std::string* err) {
    while (n > 0) {
        const ssize_t w = pwrite(fd, p, n, off);
        if (w < 0) {
            if (errno == EINTR) continue;
            *err = std::string("pwrite: ") + strerror(errno);
            return 1;
        }
        p += w; n -= (size_t)w; off += w;
    }
    return 0;
}
static int read_all(int fd, char* p, size_t n, off_t off, std::string* err) {
    while (n > 0) {
        const ssize_t r = pread(fd, p, n, off);
        if (r < 0) {
            if (errno == EINTR) continue;
            *err = std::string("pread: ") + strerror(errno);
            return 1;
        }
        if (r == 0) {
            *err = "pread: unexpected end of file (truncated restart file)";
            return 1;
        }
        p += r; n -= (size_t)r; off += r;
    }
    return 0;
}

// ---- handle-side state ----------------------------------------------------------------------------
static RestartIO* rio_get(chb_handle_s* h) {
    if (h->rio) return (RestartIO*)h->rio;
    RestartIO* r = new RestartIO();
    const char* e = getenv("CHB_IO_CHUNK_MB");
    r->chunk_bytes = (size_t)((e ? atof(e) : 64.0) * 1048576.0);
    if (r->chunk_bytes < 4096) r->chunk_bytes = 4096;
    r->chunk_bytes &= ~(size_t)15;
    bool ok = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&r->snap_done, cudaEventDisableTiming) == cudaSuccess;
    for (int b = 0; b < 2 && ok; ++b)
        ok = cudaEventCreateWithFlags(&r->copied[b], cudaEventDisableTiming) == cudaSuccess &&
             cudaMallocHost((void**)&r->pinned[b], r->chunk_bytes) == cudaSuccess;
    if (!ok) {
        chb_set_error(std::string("restart I/O setup: ") + cudaGetErrorString(cudaGetLastError()));
        delete r;
        return nullptr;
    }
    h->rio = r;
    return r;
}

static int rio_join(RestartIO* r) {
    if (!r || !r->worker_active) return 0;
    r->worker.join();
    r->worker_active = false;
    if (r->worker_rc) chb_set_error(r->worker_err);
    return r->worker_rc;
}

void chb_restart_destroy(chb_handle_s* h) {
    RestartIO* r = (RestartIO*)h->rio;
    if (!r) return;
    rio_join(r);
    for (int b = 0; b < 2; ++b) {
        if (r->pinned[b]) cudaFreeHost(r->pinned[b]);
        if (r->copied[b]) cudaEventDestroy(r->copied[b]);
    }
    if (r->snap) cudaFree(r->snap);
    if (r->snap_done) cudaEventDestroy(r->snap_done);
    if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
    delete r;
    h->rio = nullptr;
}

// drain `src` (device, file order: [c][ixl][iz][iy]) into the file: chunked D2H on the copy stream
// through two pinned buffers, chunk k written while chunk k+1 is copied
static int drain_to_file(chb_handle_s* h, RestartIO* r, const cplx* src, int fd, std::string* err) {
    const Geometry& g = h->g;
    const size_t comp_bytes = (size_t)g.nyp * g.M * sizeof(cplx);
    struct Chunk { size_t src_off; off_t file_off; size_t n; };
    auto chunk_at = [&](size_t k, Chunk* c) -> bool {   // chunks never straddle a component
        const size_t per_comp = (comp_bytes + r->chunk_bytes - 1) / r->chunk_bytes;
        if (k >= 3 * per_comp) return false;
        const size_t comp = k / per_comp, i = k % per_comp;
        const size_t o = i * r->chunk_bytes;
        c->n = (o + r->chunk_bytes <= comp_bytes) ? r->chunk_bytes : comp_bytes - o;
        c->src_off = comp * comp_bytes + o;
        c->file_off = (off_t)(chb_host_restart_offset(g.nx, g.ny, g.nz, g.nx0, (int)comp) + (long long)o);
        return true;
    };
    auto issue = [&](size_t k) -> int {
        Chunk c;
        if (!chunk_at(k, &c)) return 0;
        const int b = (int)(k & 1);
        if (cudaMemcpyAsync(r->pinned[b], reinterpret_cast<const char*>(src) + c.src_off, c.n, cudaMemcpyDeviceToHost,
                            r->copy_stream) != cudaSuccess ||
            cudaEventRecord(r->copied[b], r->copy_stream) != cudaSuccess) {
            *err = std::string("snapshot D2H: ") + cudaGetErrorString(cudaGetLastError());
            return 1;
        }
        return 0;
    };
    if (issue(0)) return 1;
    Chunk c;
    for (size_t k = 0; chunk_at(k, &c); ++k) {
        if (cudaEventSynchronize(r->copied[k & 1]) != cudaSuccess) {
            *err = std::string("snapshot D2H: ") + cudaGetErrorString(cudaGetLastError());
            return 1;
        }
        if (issue(k + 1)) return 1;     // buffer (k+1)&1 was written to the file in the previous iteration
        if (write_all(fd, r->pinned[k & 1], c.n, c.file_off, err)) return 1;
    }
    return 0;
}

extern "C" int chb_save_restart_file(chb_handle h, const char* filename, double time, int field, int async_mode) {
    if (!h || !filename) { chb_set_error("chb_save_restart_file: null argument"); return 2; }
    if (field != 0 && field != 1) { chb_set_error("chb_save_restart_file: field must be 0 (V) or 1 (F)"); return 2; }
    if (field == 1 && !h->F) { chb_set_error("chb_save_restart_file: body force not enabled"); return 2; }
    CHB_CUDA_OK(cudaSetDevice(h->device));
    RestartIO* r = rio_get(h);
    if (!r) return 1;
    if (rio_join(r)) return 1;          // one snapshot in flight at a time
    const Geometry& g = h->g;
    const size_t fld = (size_t)g.nyp * g.M;
    const auto t0 = std::chrono::steady_clock::now();

    // open / size the file; rank 0 (has_terminal) writes the header
    const int fd = open(filename, O_WRONLY | O_CREAT, 0644);
    if (fd < 0) { chb_set_error(std::string("chb_save_restart_file: open ") + filename + ": " + strerror(errno)); return 4; }
    std::string err;
    if (ftruncate(fd, (off_t)chb_host_restart_file_bytes(g.nx, g.ny, g.nz)) != 0) {
        chb_set_error(std::string("chb_save_restart_file: ftruncate: ") + strerror(errno));
        close(fd);
        return 4;
    }
    if (g.rank == 0) {
        unsigned char hdr[CHB_RESTART_HEADER_BYTES];
        chb_host_restart_header(g.nx, g.ny, g.nz, g.alfa0, g.beta0, g.ni, h->grid_a, h->grid_ymin, h->grid_ymax, time, hdr);
        if (write_all(fd, (const char*)hdr, sizeof(hdr), 0, &err)) { chb_set_error(err); close(fd); return 4; }
    }

    // (1) device-side snapshot in file order
    cplx* dst = h->P;                   // blocking mode: the products buffer is dead between time steps
    if (async_mode) {
        if (!r->snap) {
            if (cudaMalloc((void**)&r->snap, 3 * fld * sizeof(cplx)) != cudaSuccess) {
                cudaGetLastError();
                chb_set_error("chb_save_restart_file: no device memory for the asynchronous snapshot buffer (use async=0)");
                close(fd);
                return 5;
            }
            h->dev_bytes += 3 * fld * sizeof(cplx);
        }
        dst = r->snap;
    }
    const cplx* src = field == 0 ? h->V : h->F;
    cudaEvent_t e0, e1;
    CHB_CUDA_OK(cudaEventCreate(&e0));
    CHB_CUDA_OK(cudaEventCreate(&e1));
    CHB_CUDA_OK(cudaEventRecord(e0, h->stream));
    for (int c = 0; c < 3; ++c) launch_planes_to_fortran(h, src + c * fld, dst + c * fld, c, 0, g.nxB);
    CHB_CUDA_OK(cudaEventRecord(e1, h->stream));
    CHB_CUDA_OK(cudaEventRecord(r->snap_done, h->stream));
    CHB_CUDA_OK(cudaStreamWaitEvent(r->copy_stream, r->snap_done, 0));
    r->bytes = (double)(3 * fld * sizeof(cplx));

    // (2)+(3) drain
    auto finish = [h, r, dst, fd, t0, e0, e1]() -> int {
        std::string werr;
        cudaSetDevice(h->device);
        int rc = drain_to_file(h, r, dst, fd, &werr);
        if (close(fd) != 0 && !rc) { werr = std::string("close: ") + strerror(errno); rc = 1; }
        float ms = 0;
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        r->t_snapshot_ms = ms;
        r->t_total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        r->worker_rc = rc ? 4 : 0;
        r->worker_err = werr;
        return r->worker_rc;
    };
    if (async_mode) {
        r->worker_active = true;
        r->worker = std::thread(finish);
        return 0;
    }
    if (finish()) { chb_set_error(r->worker_err); return 4; }
    return 0;
}

extern "C" int chb_restart_wait(chb_handle h) {
    if (!h) { chb_set_error("chb_restart_wait: null handle"); return 2; }
    return rio_join((RestartIO*)h->rio);
}

extern "C" int chb_restart_stats(chb_handle h, double* bytes, double* snapshot_ms, double* total_s) {
    if (!h || !h->rio) { chb_set_error("chb_restart_stats: no snapshot was written"); return 2; }
    RestartIO* r = (RestartIO*)h->rio;
    if (rio_join(r)) return 1;
    if (bytes) *bytes = r->bytes;
    if (snapshot_ms) *snapshot_ms = r->t_snapshot_ms;
    if (total_s) *total_s = r->t_total_s;
    return 0;
}

extern "C" int chb_read_restart_file(chb_handle h, const char* filename, double* time) {
    if (!h || !filename) { chb_set_error("chb_read_restart_file: null argument"); return 2; }
    CHB_CUDA_OK(cudaSetDevice(h->device));
    RestartIO* r = rio_get(h);
    if (!r) return 1;
    if (rio_join(r)) return 1;
    const Geometry& g = h->g;
    const int fd = open(filename, O_RDONLY);
    if (fd < 0) {   // the reference generates an initial field instead (dnsdata.f90:705-719): the driver's job
        chb_set_error(std::string("chb_read_restart_file: cannot open ") + filename + ": " + strerror(errno));
        return 4;
    }
    std::string err;
    unsigned char hdr[CHB_RESTART_HEADER_BYTES];
    if (read_all(fd, (char*)hdr, sizeof(hdr), 0, &err)) { chb_set_error(err); close(fd); return 4; }
    int ints[3];
    double reals[7];
    memcpy(ints, hdr, 12);
    memcpy(reals, hdr + 12, 56);
    if (ints[0] != g.nx || ints[1] != g.ny || ints[2] != g.nz || reals[0] != g.alfa0 || reals[1] != g.beta0 ||
        reals[2] != g.ni || reals[3] != h->grid_a || reals[4] != h->grid_ymin || reals[5] != h->grid_ymax) {
        chb_set_error("ERROR: mismatch in metadata between restart file and dns.in. Stopping.");   // dnsdata.f90:696-703
        close(fd);
        return 3;
    }
    if (time) *time = reals[6];
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < (off_t)chb_host_restart_file_bytes(g.nx, g.ny, g.nz)) {
        chb_set_error("chb_read_restart_file: file shorter than header + 3*(nx+1)*(2nz+1)*(ny+3) complex");
        close(fd);
        return 4;
    }
    // pread chunk k into a pinned buffer while chunk k-1 crosses PCIe; stage in P (file order), then transpose
    const size_t fld = (size_t)g.nyp * g.M, comp_bytes = fld * sizeof(cplx);
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    size_t k = 0;
    for (int c = 0; c < 3; ++c) {
        const off_t base = (off_t)chb_host_restart_offset(g.nx, g.ny, g.nz, g.nx0, c);
        for (size_t o = 0; o < comp_bytes; o += r->chunk_bytes, ++k) {
            const size_t n = (o + r->chunk_bytes <= comp_bytes) ? r->chunk_bytes : comp_bytes - o;
            const int b = (int)(k & 1);
            if (k >= 2) CHB_CUDA_OK(cudaEventSynchronize(r->copied[b]));
            if (read_all(fd, r->pinned[b], n, base + (off_t)o, &err)) { chb_set_error(err); close(fd); return 4; }
            CHB_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->P) + c * comp_bytes + o, r->pinned[b], n,
                                        cudaMemcpyHostToDevice, r->copy_stream));
            CHB_CUDA_OK(cudaEventRecord(r->copied[b], r->copy_stream));
        }
    }
    close(fd);
    CHB_CUDA_OK(cudaEventRecord(r->snap_done, r->copy_stream));
    CHB_CUDA_OK(cudaStreamWaitEvent(h->stream, r->snap_done, 0));
    for (int c = 0; c < 3; ++c) launch_fortran_to_planes(h, h->P + c * fld, h->V + c * fld, c, 0, g.nxB);
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(r->copy_stream));
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}
