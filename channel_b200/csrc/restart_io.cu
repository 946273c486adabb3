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
// transpose per x-slab of a component) -- blocking mode into the work arena of the pencil transposes,
// which is dead between time steps, one slab of as many x-modes as fit after the other; asynchronous
// mode into a dedicated snapshot buffer, after which the time loop may continue at once; (2) CHB_IO_THREADS (default 4) host threads drain the snapshot, each with its own
// copy stream and two pinned buffers: thread w takes chunks w, w+NW, ... and pwrite()s chunk k while
// its next chunk is in flight over PCIe (one thread's pwrite into the page cache runs at 2-3 GB/s,
// far below PCIe, so the file side is what needs the parallelism); blocking mode joins them before
// returning, asynchronous mode leaves them running.  At 1024^3 a snapshot is 103 GB (13 GB per GPU
// on 8 GPUs): the step time hides it completely in asynchronous mode.  Reading is the mirror image.
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>

#include "../../include/channel_b200.h"
#include "../../include/channel_b200_host.h"
#include "chb_internal.h"

#define CHB_RESTART_HEADER_BYTES 68

#define CHB_IO_MAX_THREADS 8
struct IoLane {   // one host thread's copy stream, events and pinned buffers
    cudaStream_t stream = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr};
    char* pinned[2] = {nullptr, nullptr};
};
struct RestartIO {
    int nw = 1;
    IoLane lane[CHB_IO_MAX_THREADS];
    cudaEvent_t snap_done = nullptr;
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
    r->chunk_bytes = (size_t)((e ? atof(e) : 32.0) * 1048576.0);
    if (r->chunk_bytes < 4096) r->chunk_bytes = 4096;
    r->chunk_bytes &= ~(size_t)15;
    e = getenv("CHB_IO_THREADS");
    r->nw = e ? atoi(e) : 4;
    if (r->nw < 1) r->nw = 1;
    if (r->nw > CHB_IO_MAX_THREADS) r->nw = CHB_IO_MAX_THREADS;
    bool ok = cudaEventCreateWithFlags(&r->snap_done, cudaEventDisableTiming) == cudaSuccess;
    for (int w = 0; w < r->nw && ok; ++w) {
        IoLane& ln = r->lane[w];
        ok = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) == cudaSuccess;
        for (int b = 0; b < 2 && ok; ++b)
            ok = cudaEventCreateWithFlags(&ln.copied[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaMallocHost((void**)&ln.pinned[b], r->chunk_bytes) == cudaSuccess;
    }
    h->rio = r;
    if (!ok) {
        chb_set_error(std::string("restart I/O setup: ") + cudaGetErrorString(cudaGetLastError()));
        chb_restart_destroy(h);
        return nullptr;
    }
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
    for (int w = 0; w < CHB_IO_MAX_THREADS; ++w) {
        IoLane& ln = r->lane[w];
        for (int b = 0; b < 2; ++b) {
            if (ln.pinned[b]) cudaFreeHost(ln.pinned[b]);
            if (ln.copied[b]) cudaEventDestroy(ln.copied[b]);
        }
        if (ln.stream) cudaStreamDestroy(ln.stream);
    }
    if (r->snap) cudaFree(r->snap);
    if (r->snap_done) cudaEventDestroy(r->snap_done);
    delete r;
    h->rio = nullptr;
}

// A slab = the x-modes [ix0, ix0 + nix) of component c, nix*(2nz+1)*(ny+3) complex that are contiguous both in the
// file and in the device copy `dev` (file order).  Chunk k of a slab -> (offset in dev, offset in the file, bytes).
struct Slab { int c, ix0, nix; const cplx* dev; };
struct Chunk { size_t dev_off; off_t file_off; size_t n; };
static bool chunk_at(const Geometry& g, const Slab& s, size_t chunk_bytes, size_t k, Chunk* c) {
    const size_t slab_bytes = (size_t)s.nix * g.nzt * g.nyp * sizeof(cplx);
    const size_t o = k * chunk_bytes;
    if (o >= slab_bytes) return false;
    c->n = (o + chunk_bytes <= slab_bytes) ? chunk_bytes : slab_bytes - o;
    c->dev_off = o;
    c->file_off = (off_t)(chb_host_restart_offset(g.nx, g.ny, g.nz, g.nx0 + s.ix0, s.c) + (long long)o);
    return true;
}

// host thread w of nw: drain chunks w, w+nw, ... of the slab into the file; chunk k is written while this thread's
// next chunk crosses PCIe
static int drain_lane(chb_handle_s* h, RestartIO* r, int w, const Slab& s, int fd, std::string* err) {
    cudaSetDevice(h->device);
    IoLane& ln = r->lane[w];
    const size_t nw = (size_t)r->nw;
    auto issue = [&](size_t i) -> int {   // i-th chunk of this lane
        Chunk c;
        if (!chunk_at(h->g, s, r->chunk_bytes, w + i * nw, &c)) return 0;
        const int b = (int)(i & 1);
        if (cudaMemcpyAsync(ln.pinned[b], reinterpret_cast<const char*>(s.dev) + c.dev_off, c.n, cudaMemcpyDeviceToHost,
                            ln.stream) != cudaSuccess ||
            cudaEventRecord(ln.copied[b], ln.stream) != cudaSuccess) {
            *err = std::string("snapshot D2H: ") + cudaGetErrorString(cudaGetLastError());
            return 1;
        }
        return 0;
    };
    if (cudaStreamWaitEvent(ln.stream, r->snap_done, 0) != cudaSuccess) { *err = "snapshot: cudaStreamWaitEvent failed"; return 1; }
    if (issue(0)) return 1;
    Chunk c;
    for (size_t i = 0; chunk_at(h->g, s, r->chunk_bytes, w + i * nw, &c); ++i) {
        if (cudaEventSynchronize(ln.copied[i & 1]) != cudaSuccess) {
            *err = std::string("snapshot D2H: ") + cudaGetErrorString(cudaGetLastError());
            return 1;
        }
        if (issue(i + 1)) return 1;     // buffer (i+1)&1 went to the file in the previous iteration
        if (write_all(fd, ln.pinned[i & 1], c.n, c.file_off, err)) return 1;
    }
    return 0;
}

// all lanes: nw-1 extra threads + the calling one
static int drain_to_file(chb_handle_s* h, RestartIO* r, const Slab& s, int fd, std::string* err) {
    std::string errs[CHB_IO_MAX_THREADS];
    int rcs[CHB_IO_MAX_THREADS] = {0};
    std::thread th[CHB_IO_MAX_THREADS];
    for (int w = 1; w < r->nw; ++w) th[w] = std::thread([&, w]() { rcs[w] = drain_lane(h, r, w, s, fd, &errs[w]); });
    rcs[0] = drain_lane(h, r, 0, s, fd, &errs[0]);
    int rc = 0;
    for (int w = 0; w < r->nw; ++w) {
        if (w) th[w].join();
        if (rcs[w] && !rc) { rc = rcs[w]; *err = errs[w]; }
    }
    return rc;
}

// x-modes per slab when the work arena is the staging area
static int arena_slab_width(const chb_handle_s* h) {
    const Geometry& g = h->g;
    const size_t per_ix = (size_t)g.nzt * g.nyp * sizeof(cplx);
    size_t n = (h->arena_bytes - h->stage_off) / per_ix;
    if (n > (size_t)g.nxB) n = g.nxB;
    return (int)n;
}

// closes the file and destroys the events on every path out of chb_save_restart_file
struct SaveGuard {
    int fd = -1;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~SaveGuard() {
        if (fd >= 0) close(fd);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

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
    if (!async_mode && arena_slab_width(h) < 1) {
        chb_set_error("chb_save_restart_file: work arena smaller than one x-mode of the field (raise CHB_WORK_GB)");
        return 5;
    }

    // open / size the file; rank 0 (has_terminal) writes the header
    auto guard = std::make_shared<SaveGuard>();
    guard->fd = open(filename, O_WRONLY | O_CREAT, 0644);
    const int fd = guard->fd;
    if (fd < 0) { chb_set_error(std::string("chb_save_restart_file: open ") + filename + ": " + strerror(errno)); return 4; }
    std::string err;
    if (ftruncate(fd, (off_t)chb_host_restart_file_bytes(g.nx, g.ny, g.nz)) != 0) {
        chb_set_error(std::string("chb_save_restart_file: ftruncate: ") + strerror(errno));
        return 4;
    }
    if (g.rank == 0) {
        unsigned char hdr[CHB_RESTART_HEADER_BYTES];
        chb_host_restart_header(g.nx, g.ny, g.nz, g.alfa0, g.beta0, g.ni, h->grid_a, h->grid_ymin, h->grid_ymax, time, hdr);
        if (write_all(fd, (const char*)hdr, sizeof(hdr), 0, &err)) { chb_set_error(err); return 4; }
    }
    const cplx* src = field == 0 ? h->V : h->F;
    CHB_CUDA_OK(cudaEventCreate(&guard->e0));
    CHB_CUDA_OK(cudaEventCreate(&guard->e1));
    r->bytes = (double)(3 * fld * sizeof(cplx));

    if (async_mode) {
        // (1) the whole field in file order into the private snapshot buffer, (2)+(3) drained by a worker thread
        if (!r->snap) {
            if (cudaMalloc((void**)&r->snap, 3 * fld * sizeof(cplx)) != cudaSuccess) {
                cudaGetLastError();
                chb_set_error("chb_save_restart_file: no device memory for the asynchronous snapshot buffer (use async=0)");
                return 5;
            }
            h->dev_bytes += 3 * fld * sizeof(cplx);
        }
        CHB_CUDA_OK(cudaEventRecord(guard->e0, h->stream));
        for (int c = 0; c < 3; ++c) launch_planes_to_fortran(h, src + c * fld, r->snap + c * fld, 0, g.nxB, h->stream);
        CHB_CUDA_OK(cudaEventRecord(guard->e1, h->stream));
        CHB_CUDA_OK(cudaEventRecord(r->snap_done, h->stream));
        auto finish = [h, r, guard, t0]() -> int {
            std::string werr;
            cudaSetDevice(h->device);
            int rc = 0;
            for (int c = 0; c < 3 && !rc; ++c) {
                const Slab s = {c, 0, h->g.nxB, r->snap + (size_t)c * h->g.nyp * h->g.M};
                rc = drain_to_file(h, r, s, guard->fd, &werr);
            }
            if (close(guard->fd) != 0 && !rc) { werr = std::string("close: ") + strerror(errno); rc = 1; }
            guard->fd = -1;
            float ms = 0;
            cudaEventSynchronize(guard->e1);
            cudaEventElapsedTime(&ms, guard->e0, guard->e1);
            r->t_snapshot_ms = ms;
            r->t_total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            r->worker_rc = rc ? 4 : 0;
            r->worker_err = werr;
            return r->worker_rc;
        };
        r->worker_active = true;
        r->worker = std::thread(finish);
        return 0;
    }

    // blocking mode: slab by slab through the work arena (dead between time steps)
    cplx* stage = reinterpret_cast<cplx*>(h->arena + h->stage_off);
    const int wmax = arena_slab_width(h);
    double t_ms = 0;
    for (int c = 0; c < 3; ++c) {
        for (int ix0 = 0; ix0 < g.nxB; ix0 += wmax) {
            const int nix = (ix0 + wmax <= g.nxB) ? wmax : g.nxB - ix0;
            CHB_CUDA_OK(cudaEventRecord(guard->e0, h->stream));
            launch_planes_to_fortran(h, src + c * fld, stage, ix0, nix, h->stream);
            CHB_CUDA_OK(cudaEventRecord(guard->e1, h->stream));
            CHB_CUDA_OK(cudaEventRecord(r->snap_done, h->stream));
            const Slab s = {c, ix0, nix, stage};
            if (drain_to_file(h, r, s, fd, &err)) { chb_set_error(err); return 4; }
            float ms = 0;
            CHB_CUDA_OK(cudaEventElapsedTime(&ms, guard->e0, guard->e1));
            t_ms += ms;
        }
    }
    guard->fd = -1;
    if (close(fd) != 0) { chb_set_error(std::string("close: ") + strerror(errno)); return 4; }
    r->t_snapshot_ms = t_ms;
    r->t_total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    r->worker_rc = 0;
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

struct FdGuard {
    int fd;
    ~FdGuard() { if (fd >= 0) close(fd); }
};

extern "C" int chb_read_restart_file(chb_handle h, const char* filename, double* time) {
    if (!h || !filename) { chb_set_error("chb_read_restart_file: null argument"); return 2; }
    CHB_CUDA_OK(cudaSetDevice(h->device));
    RestartIO* r = rio_get(h);
    if (!r) return 1;
    if (rio_join(r)) return 1;
    const Geometry& g = h->g;
    FdGuard fg = {open(filename, O_RDONLY)};
    const int fd = fg.fd;
    if (fd < 0) {   // the reference generates an initial field instead (dnsdata.f90:705-719): the driver's job
        chb_set_error(std::string("chb_read_restart_file: cannot open ") + filename + ": " + strerror(errno));
        return 4;
    }
    std::string err;
    unsigned char hdr[CHB_RESTART_HEADER_BYTES];
    if (read_all(fd, (char*)hdr, sizeof(hdr), 0, &err)) { chb_set_error(err); return 4; }
    int ints[3];
    double reals[7];
    memcpy(ints, hdr, 12);
    memcpy(reals, hdr + 12, 56);
    if (ints[0] != g.nx || ints[1] != g.ny || ints[2] != g.nz || reals[0] != g.alfa0 || reals[1] != g.beta0 ||
        reals[2] != g.ni || reals[3] != h->grid_a || reals[4] != h->grid_ymin || reals[5] != h->grid_ymax) {
        chb_set_error("ERROR: mismatch in metadata between restart file and dns.in. Stopping.");   // dnsdata.f90:696-703
        return 3;
    }
    if (time) *time = reals[6];
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < (off_t)chb_host_restart_file_bytes(g.nx, g.ny, g.nz)) {
        chb_set_error("chb_read_restart_file: file shorter than header + 3*(nx+1)*(2nz+1)*(ny+3) complex");
        return 4;
    }
    // slab by slab through the work arena: thread w preads its chunk i into a pinned buffer while its chunk i-1 crosses
    // PCIe; the slab, complete in file order, is then transposed into the device layout
    const size_t fld = (size_t)g.nyp * g.M;
    const int wmax = arena_slab_width(h);
    if (wmax < 1) { chb_set_error("chb_read_restart_file: work arena smaller than one x-mode of the field (raise CHB_WORK_GB)"); return 5; }
    cplx* stage = reinterpret_cast<cplx*>(h->arena + h->stage_off);
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < 3; ++c) {
        for (int ix0 = 0; ix0 < g.nxB; ix0 += wmax) {
            const int nix = (ix0 + wmax <= g.nxB) ? wmax : g.nxB - ix0;
            const Slab s = {c, ix0, nix, stage};
            std::string errs[CHB_IO_MAX_THREADS];
            int rcs[CHB_IO_MAX_THREADS] = {0};
            std::thread th[CHB_IO_MAX_THREADS];
            auto load_lane = [&](int w) {
                cudaSetDevice(h->device);
                IoLane& ln = r->lane[w];
                Chunk ck;
                for (size_t i = 0; chunk_at(g, s, r->chunk_bytes, w + i * (size_t)r->nw, &ck); ++i) {
                    const int b = (int)(i & 1);
                    if (i >= 2 && cudaEventSynchronize(ln.copied[b]) != cudaSuccess) { rcs[w] = 1; errs[w] = "restart H2D failed"; return; }
                    if (read_all(fd, ln.pinned[b], ck.n, ck.file_off, &errs[w])) { rcs[w] = 4; return; }
                    if (cudaMemcpyAsync(reinterpret_cast<char*>(stage) + ck.dev_off, ln.pinned[b], ck.n, cudaMemcpyHostToDevice,
                                        ln.stream) != cudaSuccess ||
                        cudaEventRecord(ln.copied[b], ln.stream) != cudaSuccess) { rcs[w] = 1; errs[w] = "restart H2D failed"; return; }
                }
                if (cudaStreamSynchronize(ln.stream) != cudaSuccess) { rcs[w] = 1; errs[w] = "restart H2D failed"; }
            };
            for (int w = 1; w < r->nw; ++w) th[w] = std::thread(load_lane, w);
            load_lane(0);
            for (int w = 1; w < r->nw; ++w) th[w].join();
            for (int w = 0; w < r->nw; ++w)
                if (rcs[w]) { chb_set_error("chb_read_restart_file: " + errs[w]); return rcs[w]; }
            launch_fortran_to_planes(h, stage, h->V + c * fld, ix0, nix, h->stream);
            CHB_CUDA_OK(cudaStreamSynchronize(h->stream));   // the arena is reused by the next slab
        }
    }
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}
