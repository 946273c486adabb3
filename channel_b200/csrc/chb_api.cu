// chb_api.cu - the extern "C" entry points declared in include/channel_b200.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/channel_b200.h"
#include "chb_internal.h"

static thread_local std::string g_err;
void chb_set_error(const std::string& s) { g_err = s; }
extern "C" const char* chb_last_error(void) { return g_err.c_str(); }
extern "C" int chb_version(void) { return 2000; }   // 2000: RHS in place in V, products per chunk (chb_debug_capture_products), barrier timeouts

#define CHB_REQUIRE(cond, msg)  \
    do {                        \
        if (!(cond)) {          \
            chb_set_error(msg); \
            return 2;           \
        }                       \
    } while (0)

// ---- timing -----------------------------------------------------------------------------
ScopedKernelTimer::ScopedKernelTimer(chb_handle_s* h_, const char* name_, cudaStream_t st_)
    : h(h_), name(name_), on(h_->timer.on), st(st_ ? st_ : h_->stream) {
    if (on) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
}
ScopedKernelTimer::~ScopedKernelTimer() {
    if (on) {
        cudaEventRecord(e1, st);
        h->timer.pending.push_back({name, {e0, e1}});
    }
}
void chb_timer_flush(chb_handle_s* h) {
    if (h->timer.pending.empty()) return;
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->side_stream);
    if (h->sA && h->sA != h->stream) cudaStreamSynchronize(h->sA);
    if (h->sB && h->sB != h->stream) cudaStreamSynchronize(h->sB);
    for (auto& p : h->timer.pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, p.second.first, p.second.second);
        auto& r = h->timer.recs[p.first];
        r.ms += ms;
        r.n += 1;
        cudaEventDestroy(p.second.first);
        cudaEventDestroy(p.second.second);
    }
    h->timer.pending.clear();
}

// ---- FFT plans ----------------------------------------------------------------------------
// radices for n = 2^a 3^b: the 3 first (largest stride), then 8s, then a trailing 4 or 2
static bool build_plan(int n, FftPlan* pl, std::vector<int>* rev) {
    pl->n = n;
    pl->npass = 0;
    int r = n;
    if (r % 3 == 0) {
        pl->radix[pl->npass++] = 3;
        r /= 3;
    }
    if (r % 3 == 0) return false;
    int a = 0;
    while (r % 2 == 0) {
        r /= 2;
        ++a;
    }
    if (r != 1) return false;
    while (a >= 3) {
        pl->radix[pl->npass++] = 8;
        a -= 3;
    }
    if (a == 2) pl->radix[pl->npass++] = 4;
    if (a == 1) pl->radix[pl->npass++] = 2;
    if (pl->npass > CHB_MAX_PASSES) return false;
    if (rev) {
        rev->resize(n);
        for (int k = 0; k < n; ++k) {
            int kk = k, pos = 0, stride = n;
            for (int t = 0; t < pl->npass; ++t) {
                const int R = pl->radix[t];
                stride /= R;
                pos += (kk % R) * stride;
                kk /= R;
            }
            (*rev)[k] = pos;
        }
    }
    return true;
}

static std::vector<double> twiddles(int n, int count, int denom) {
    std::vector<double> w(2 * (size_t)count);
    for (int e = 0; e < count; ++e) {
        const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / (long double)denom;
        w[2 * e] = (double)cosl(a);
        w[2 * e + 1] = (double)sinl(a);
    }
    (void)n;
    return w;
}

static thread_local size_t g_alloc_bytes = 0;
template <typename T>
static int dev_alloc(T** p, size_t count) {
    CHB_CUDA_OK(cudaMalloc((void**)p, count * sizeof(T)));
    g_alloc_bytes += count * sizeof(T);
    CHB_CUDA_OK(cudaMemset(*p, 0, count * sizeof(T)));
    return 0;
}
template <typename T>
static int dev_upload(T** p, const std::vector<T>& v) {
    CHB_CUDA_OK(cudaMalloc((void**)p, v.size() * sizeof(T)));
    CHB_CUDA_OK(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// ---- multi-GPU (transpose.cu) ----
int chb_nccl_init(chb_handle_s* h, const char* id);
void chb_nccl_destroy(chb_handle_s* h);
int chb_nccl_unique_id(char* id);
int chb_allreduce_max_cfl(chb_handle_s* h);
int chb_bcast_scalars(chb_handle_s* h);

extern "C" int chb_get_nccl_unique_id(char* id) { return chb_nccl_unique_id(id); }

void chb_select_lane(chb_handle_s* h, int L) {
    const Lane& ln = h->lane[L];
    h->cur_lane = L;
    h->A = ln.A; h->Ar = ln.Ar; h->B = ln.B; h->Br = ln.Br; h->Pc = ln.Pc;
    h->Aw = ln.Aw; h->Bw = ln.Bw;
    h->flags = ln.flags;
    for (int q = 0; q < CHB_MAX_RANKS; ++q) h->peer_flags[q] = ln.peer_flags[q];
}

// Layout of the work arena: [0, 4096) barrier flags of every lane + the barrier error word; then the velocity receive
// buffers Ar of all lanes; then, per lane, Br, Pc (+ A, B in NCCL mode); everything aligned to 256 bytes.  Identical on
// every rank (np is agreed on), so a peer's buffers are its arena base + these offsets.
// The part behind the Ar buffers doubles as the staging area of the host <-> device field transfers and of the restart
// files.  Ar is excluded on purpose: between two sweeps a peer that is ahead may already run zfwd of its next sweep,
// whose stores land in this rank's Ar; Br is only written behind a barrier this rank takes part in (after its own
// staging work, in stream order), Pc / A / B are private.
#define CHB_ARENA_HEAD 4096
struct ArenaLayout {
    size_t ar_bytes;           // one lane's Ar
    size_t br, pc, a, b;       // offsets inside a lane's private part
    size_t lane_bytes;         // private part of one lane
    size_t stage_off, total;
};
static ArenaLayout arena_layout(const Geometry& g, size_t np, int nlanes, bool nccl_mode) {
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t na = (size_t)3 * np * g.nzd * g.nxB * sizeof(cplx), nb = 2 * na, npc = (size_t)6 * np * g.M * sizeof(cplx);
    ArenaLayout L;
    L.ar_bytes = up(na);
    size_t o = 0;
    L.br = o; o = up(o + nb);
    L.pc = o; o = up(o + npc);
    L.a = L.b = 0;
    if (nccl_mode) {
        L.a = o; o = up(o + na);
        L.b = o; o = up(o + nb);
    }
    L.lane_bytes = o;
    L.stage_off = CHB_ARENA_HEAD + (size_t)nlanes * L.ar_bytes;
    L.total = L.stage_off + (size_t)nlanes * o;
    return L;
}

// ---- create / destroy ---------------------------------------------------------------------
static int create_impl(chb_handle_s* h, int nx, int ny, int nz, int nxd, int nzd, double alfa0, double beta0, double ni,
                       double a, double ymin, double ymax, int rank, int nranks, const char* nccl_id, int device);

extern "C" int chb_create(chb_handle* out, int nx, int ny, int nz, int nxd, int nzd, double alfa0, double beta0,
                          double ni, double a, double ymin, double ymax, int rank, int nranks, const char* nccl_id,
                          int device) {
    CHB_REQUIRE(out != nullptr, "chb_create: null handle pointer");
    CHB_REQUIRE(nx >= 1 && ny >= 8 && nz >= 1, "chb_create: need nx>=1, ny>=8, nz>=1");
    CHB_REQUIRE(nxd >= nx + 1 && nxd % 2 == 0, "chb_create: nxd must be even and >= nx+1");
    CHB_REQUIRE(nzd >= 2 * nz + 1, "chb_create: nzd must be >= 2nz+1");
    CHB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "chb_create: bad rank/nranks");
    CHB_REQUIRE(nranks <= CHB_MAX_RANKS, "chb_create: at most 8 ranks (one NVSwitch node)");
    CHB_REQUIRE((nx + 1) % nranks == 0 && nzd % nranks == 0,
                "chb_create: nranks must divide nx+1 and nzd (README.md:154)");
    CHB_REQUIRE(nranks == 1 || nccl_id != nullptr, "chb_create: nccl_id required when nranks>1");
    CHB_REQUIRE(mean_mode_smem_doubles(ny) * sizeof(double) <= 227 * 1024,
                "chb_create: ny too large for the mean-mode kernel (its column lives in shared memory: ny <= 3056)");
    int ndev = 0;
    CHB_CUDA_OK(cudaGetDeviceCount(&ndev));
    CHB_REQUIRE(ndev > 0, "chb_create: no CUDA device (this library has no CPU fallback)");
    CHB_CUDA_OK(cudaSetDevice(device));
    chb_handle_s* h = new chb_handle_s();   // value-initialised: every pointer, stream and event starts out null
    h->device = device;
    const int rc = create_impl(h, nx, ny, nz, nxd, nzd, alfa0, beta0, ni, a, ymin, ymax, rank, nranks, nccl_id, device);
    if (rc) {   // release whatever was allocated before the failure; keep the error text of the failure
        const std::string err = g_err;
        chb_destroy(h);
        g_err = err;
        *out = nullptr;
        return rc;
    }
    *out = h;
    return 0;
}

static int create_impl(chb_handle_s* h, int nx, int ny, int nz, int nxd, int nzd, double alfa0, double beta0, double ni,
                       double a, double ymin, double ymax, int rank, int nranks, const char* nccl_id, int device) {
    g_alloc_bytes = 0;
    h->sw0 = h->sw1 = nullptr;
    memset(&h->g, 0, sizeof(h->g));
    Geometry& g = h->g;
    g.nx = nx; g.ny = ny; g.nz = nz; g.nxd = nxd; g.nzd = nzd;
    g.nyp = ny + 3; g.nzt = 2 * nz + 1;
    g.rank = rank; g.nranks = nranks;
    // mpi_transpose.f90:214-215
    chb_decompose(nx + 1, nzd, nranks, rank, &g.nx0, &g.nxN, &g.nz0, &g.nzN);
    g.nxB = g.nxN - g.nx0 + 1;
    g.nzB = g.nzN - g.nz0 + 1;
    g.M = (long long)g.nxB * g.nzt;
    g.alfa0 = alfa0; g.beta0 = beta0; g.ni = ni;
    const double PI = 3.1415926535897932384626433832795028841971;  // dnsdata.f90:26
    g.dx = PI / (alfa0 * nxd); g.dz = 2.0 * PI / (beta0 * nzd); g.factor = 1.0 / (2.0 * nxd * nzd);  // :124
    h->device = device;
    h->rio = nullptr;
    h->grid_a = a; h->grid_ymin = ymin; h->grid_ymax = ymax;
    {
        // lines per z-pass CTA.  zbwd: 2 for the long lines (4 CTAs/SM at nzd = 1536), else 4.  zfwd: 4
        // (a warp's stores into the tiled velocity buffer are 512 contiguous bytes for any value).
        const char* e = getenv("CHB_ZB_LPC");
        h->zb_lines_per_cta = e ? atoi(e) : (nzd >= 1536 ? 2 : 4);
        e = getenv("CHB_ZF_LPC");
        h->zf_lines_per_cta = e ? atoi(e) : (nzd >= 3072 ? 2 : 4);
        e = getenv("CHB_FFT3");
        h->use_fft3 = (e && atoi(e) == 0) ? 0 : 1;
        e = getenv("CHB_ZF_DIRECT");
        h->zf_direct = e ? atoi(e) : 0;
        e = getenv("CHB_Z_TPL");
        h->z_tpl = e ? atoi(e) : 0;
        e = getenv("CHB_SOLVE_PF");
        h->solve_pf = e ? atoi(e) : 0;
        h->rhs_state = nullptr;
        // two threads per innermost butterfly position of the x-pass: measured 6 % faster at nxd = 1536 (one CTA per
        // SM either way), 34 % slower at nxd = 768 (profiles/r2a_variants.md)
        e = getenv("CHB_XPASS_SPLIT");
        h->xpass_split = e ? atoi(e) : (nxd == 1536 ? 1 : 0);
        // persistent x-pass with prefetched inputs: measured 16 % faster at nxd = 1536 on top of the split (14.0 against
        // 16.6 ms), 14 % slower at nxd = 768 where it costs the third CTA per SM (profiles/r2c_r2e_single_gpu.md)
        e = getenv("CHB_XPASS_PERSIST");
        h->xpass_persist = e ? atoi(e) : (nxd == 1536 ? 1 : 0);
        // x tiles of the work buffers (transpose_index.h): products 8 wide (128-byte store segments in
        // the x-pass), velocities as wide as the lines of one zfwd CTA
        g.tw = (g.nxB % 8 == 0) ? 3 : ((g.nxB % 4 == 0) ? 2 : 0);
        e = getenv("CHB_TW");
        if (e && (g.nxB % (1 << atoi(e)) == 0)) g.tw = atoi(e);
        // single GPU: row-major (the x-pass reads whole 128-byte lines); multi GPU: tiled, so that the
        // warps of zfwd store 512 contiguous bytes into peer HBM over NVLink
        g.twa = -1;
        if (nranks > 1)
            for (int t = 1; t <= 3; ++t)
                if ((1 << t) == h->zf_lines_per_cta && g.nxB % (1 << t) == 0) g.twa = t;
        e = getenv("CHB_TWA");
        if (nranks > 1 && e && (atoi(e) < 0 || g.nxB % (1 << atoi(e)) == 0)) g.twa = atoi(e);
    }
    h->launches = 0;
    h->tables_set = false;
    h->F = nullptr;
    h->nccl_comm = nullptr;
    memset(&h->bf, 0, sizeof(h->bf));
    CHB_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CHB_CUDA_OK(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    CHB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CHB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));

    std::vector<int> rev;
    CHB_REQUIRE(build_plan(nzd, &h->plan_z, &rev), "chb_create: nzd must be 2^a or 3*2^a (fftFIT, ffts.f90:78-86)");
    CHB_REQUIRE(build_plan(nxd, &h->plan_x, nullptr), "chb_create: nxd must be 2^a or 3*2^a (fftFIT, ffts.f90:78-86)");
    if (dev_upload(&h->rev_z, rev)) return 1;
    {
        std::vector<double> wz = twiddles(nzd, nzd, nzd), wx = twiddles(nxd, nxd, nxd), wh = twiddles(nxd, nxd, 2 * nxd);
        double *pz, *px, *ph;
        if (dev_upload(&pz, wz) || dev_upload(&px, wx) || dev_upload(&ph, wh)) return 1;
        h->Wz = (cplx*)pz; h->Wx = (cplx*)px; h->Wh = (cplx*)ph;
    }
    const size_t fld = (size_t)g.nyp * g.M;
    // resident state: V 3 + oldrhs 2 + UL checkpoints 0.5 complex per point (the RHS is written in place into V, the
    // spectral products live per chunk in the arena): the reference's footprint, dnsdata.f90:144-146,667-671
    if (dev_alloc(&h->V, 3 * fld) || dev_alloc(&h->oldrhs, 2 * fld) ||
        dev_alloc(&h->ckpt, (size_t)((ny - 1) / CHB_SOLVE_K + 1) * 8 * g.M) || dev_alloc(&h->rhs_state, (size_t)32 * g.M))
        return 1;
    if (dev_alloc(&h->t_y, (size_t)g.nyp) || dev_alloc(&h->t_dy, (size_t)g.nyp) ||
        dev_alloc(&h->t_d0, (size_t)g.nyp * 5) || dev_alloc(&h->t_d1, (size_t)g.nyp * 5) ||
        dev_alloc(&h->t_d2, (size_t)g.nyp * 5) || dev_alloc(&h->t_d4, (size_t)g.nyp * 5) ||
        dev_alloc(&h->t_D0mat, (size_t)(ny + 1) * 5) || dev_alloc(&h->t_rows, (size_t)g.nyp * 25))
        return 1;
    if (dev_alloc(&h->sc, 1)) return 1;
    CHB_CUDA_OK(cudaMallocHost((void**)&h->sc_host, sizeof(DevScalars)));
    memset(h->sc_host, 0, sizeof(DevScalars));
    if (nranks > 1 && chb_nccl_init(h, nccl_id)) return 1;
    // ---- work arena: chunks of planes of the pencil transposes, one or two lanes ----
    {
        const char* e = getenv("CHB_P2P");
        h->p2p = (nranks > 1 && !(e && atoi(e) == 0)) ? 1 : 0;
        // two lanes: the kernels that carry the transposes run on their own stream (and SM partition) one chunk ahead of
        // the z-passes; default from 4 GPUs on for the large transforms
        // Measured (profiles/r2_multi_gpu.md): the pipeline wins 11-12 % on the headline grid at 4 and 8 GPUs (nxd = 1536:
        // the persistent x-pass is one CTA per SM and scales with its SM count) and 4 % on config 3 at 8 GPUs; at
        // nxd = 768 on 4 GPUs it loses to the sequential sweep (three x-pass CTAs per SM hide each other's phases;
        // on fewer SMs the kernel loses more than its share), as it does at 2 GPUs (little of the step is
        // NVLink-bound).  The NCCL fallback always runs one lane.
        const bool nccl_mode = nranks > 1 && !h->p2p;
        e = getenv("CHB_LANES");
        h->nlanes = e ? (atoi(e) == 2 ? 2 : 1) : ((nranks >= 8 || (nranks >= 4 && nxd >= 1536)) ? 2 : 1);
        if (nccl_mode) h->nlanes = 1;
        // budget: CHB_WORK_GB (default 10 GB, at most a quarter of the free device memory); larger chunks mean fewer
        // launches, fewer partially filled last waves and fewer carried-accumulator round trips of the RHS assembly
        size_t free_b = 0, total_b = 0;
        CHB_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
        e = getenv("CHB_WORK_GB");
        size_t budget = (size_t)((e ? atof(e) : 10.0) * 1073741824.0);
        if (budget > free_b / 4) budget = free_b / 4;
        const ArenaLayout AL1 = arena_layout(g, 1, 1, nccl_mode);
        const size_t per_plane = AL1.ar_bytes + AL1.lane_bytes;
        long long np = (long long)(budget / h->nlanes / per_plane);
        if (np < 1) np = 1;
        // with two lanes at least four chunks per sweep, so that the pipeline has something to overlap
        const long long cap = h->nlanes == 2 ? (g.nyp + 3) / 4 : g.nyp;
        if (np > cap) np = cap;
        // every rank must use the same chunk size (it is part of the peer-visible buffer layout and fixes the number of
        // barriers per sweep) and the same mode switches: agree on the minimum, refuse differing switches
        if (nranks > 1) {
            long long v[5] = {np, h->nlanes, -(long long)h->nlanes, h->p2p, -(long long)h->p2p};
            if (chb_allreduce_min_i64(h, v, 5)) return 1;
            CHB_REQUIRE(v[1] == -v[2] && v[3] == -v[4], "chb_create: CHB_LANES / CHB_P2P differ between the ranks");
            np = v[0];
        }
        h->chunk_planes = (int)np;
        const ArenaLayout AL = arena_layout(g, (size_t)np, h->nlanes, nccl_mode);
        h->arena_bytes = AL.total;
        h->stage_off = AL.stage_off;
        CHB_CUDA_OK(cudaMalloc((void**)&h->arena, AL.total));
        g_alloc_bytes += AL.total;
        CHB_CUDA_OK(cudaMemset(h->arena, 0, AL.total));
        h->p2p_error = reinterpret_cast<unsigned long long*>(h->arena + 2048);
        h->n_ipc_opened = 0;
        const size_t na = (size_t)3 * np * nzd * g.nxB, nb = 2 * na;
        for (int L = 0; L < h->nlanes; ++L) {
            Lane& ln = h->lane[L];
            memset(&ln, 0, sizeof(ln));
            CHB_CUDA_OK(cudaEventCreateWithFlags(&ln.evA, cudaEventDisableTiming));
            CHB_CUDA_OK(cudaEventCreateWithFlags(&ln.evZ, cudaEventDisableTiming));
            char* base = h->arena + AL.stage_off + (size_t)L * AL.lane_bytes;
            ln.Ar = reinterpret_cast<cplx*>(h->arena + CHB_ARENA_HEAD + (size_t)L * AL.ar_bytes);
            ln.Br = reinterpret_cast<cplx*>(base + AL.br);
            ln.Pc = reinterpret_cast<cplx*>(base + AL.pc);
            if (nccl_mode) {
                ln.A = reinterpret_cast<cplx*>(base + AL.a);
                ln.B = reinterpret_cast<cplx*>(base + AL.b);
            }
            ln.flags = reinterpret_cast<unsigned long long*>(h->arena) + (size_t)L * CHB_MAX_RANKS;
            // pack-side store targets (PeerPtrs): element for peer q at p[q] + index(block = rank, ...)
            const ptrdiff_t blkA = (ptrdiff_t)(na / nranks), blkB = (ptrdiff_t)(nb / nranks);
            for (int q = 0; q < nranks; ++q) {
                ln.Aw.p[q] = (nccl_mode ? ln.A : ln.Ar) + (ptrdiff_t)(q - rank) * blkA;
                ln.Bw.p[q] = (nccl_mode ? ln.B : ln.Br) + (ptrdiff_t)(q - rank) * blkB;
            }
        }
        chb_select_lane(h, 0);
        // streams of the chunk pipeline
        h->sA = h->sB = h->stream;
        h->green[0] = h->green[1] = nullptr;
        h->green_sms[0] = h->green_sms[1] = 0;
        if (h->nlanes == 2) {
            // CHB_GREEN=<SMs of the x-pass partition> (default: 54 % of the SMs on several GPUs - 80 + 68 measured
            // best of 52 / 66 / 80 at 4 GPUs and of 64 / 72 / 80 at 8 -, off on one): the two streams get disjoint SM partitions through CUDA
            // green contexts, so that the z-passes really run beside the x-pass instead of behind it; plain streams
            // if the driver refuses
            e = getenv("CHB_GREEN");
            int sms_a = e ? atoi(e) : (nranks > 1 ? -1 : 0);
            if (sms_a != 0 && chb_green_create(h, sms_a) != 0) sms_a = 0;
            if (sms_a == 0) {
                CHB_CUDA_OK(cudaStreamCreateWithFlags(&h->sA, cudaStreamNonBlocking));
                CHB_CUDA_OK(cudaStreamCreateWithFlags(&h->sB, cudaStreamNonBlocking));
            }
        }
        h->cstream = h->sA;
    }
    if (h->p2p && chb_p2p_setup(h)) return 1;
    CHB_CUDA_OK(cudaDeviceSynchronize());
    h->dev_bytes = g_alloc_bytes;
    return 0;
}

extern "C" int chb_destroy(chb_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    chb_timer_flush(h);
    chb_restart_destroy(h);
    chb_nccl_destroy(h);
    cudaFree(h->V); cudaFree(h->oldrhs); cudaFree(h->ckpt);
    if (h->F) cudaFree(h->F);
    if (h->rhs_state) cudaFree(h->rhs_state);
    if (h->P_dbg) cudaFree(h->P_dbg);
    if (h->cv_Vold) cudaFree(h->cv_Vold);
    if (h->cv_uconv) cudaFree(h->cv_uconv);
    if (h->p2p) chb_p2p_teardown(h);
    for (int L = 0; L < CHB_MAX_LANES; ++L) {
        Lane& ln = h->lane[L];
        if (ln.evA) cudaEventDestroy(ln.evA);
        if (ln.evZ) cudaEventDestroy(ln.evZ);
    }
    if (h->sA && h->sA != h->stream) cudaStreamDestroy(h->sA);
    if (h->sB && h->sB != h->stream) cudaStreamDestroy(h->sB);
    chb_green_destroy(h);
    if (h->arena) cudaFree(h->arena);
    cudaFree(h->Wz); cudaFree(h->Wx); cudaFree(h->Wh); cudaFree(h->rev_z);
    cudaFree(h->t_y); cudaFree(h->t_dy); cudaFree(h->t_d0); cudaFree(h->t_d1); cudaFree(h->t_d2); cudaFree(h->t_d4);
    cudaFree(h->t_D0mat); cudaFree(h->t_rows); cudaFree(h->sc);
    if (h->bf.mask_y) cudaFree(h->bf.mask_y);
    if (h->bf.mask_z) cudaFree(h->bf.mask_z);
    if (h->bf.mask_yz) cudaFree(h->bf.mask_yz);
    if (h->sc_host) cudaFreeHost(h->sc_host);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->sw0) cudaEventDestroy(h->sw0);
    if (h->sw1) cudaEventDestroy(h->sw1);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();   // a handle torn down half-built must not leave an error behind for the next call
    delete h;
    return 0;
}

extern "C" int chb_get_decomposition(chb_handle h, int* nx0, int* nxN, int* nz0, int* nzN) {
    CHB_REQUIRE(h, "null handle");
    *nx0 = h->g.nx0; *nxN = h->g.nxN; *nz0 = h->g.nz0; *nzN = h->g.nzN;
    return 0;
}

// ---- tables --------------------------------------------------------------------------------
extern "C" int chb_set_tables(chb_handle h, const double* y, const double* d0, const double* d1, const double* d2,
                              const double* d4, const double* d140, const double* d14m1, const double* d240,
                              const double* d24m1, const double* d14n, const double* d14np1, const double* d24n,
                              const double* d24np1, const double* v0bc, const double* v0m1bc, const double* vnbc,
                              const double* vnp1bc, const double* eta0bc, const double* eta0m1bc,
                              const double* etanbc, const double* etanp1bc, const double* D0mat) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    const Geometry& g = h->g;
    const int ny = g.ny;
    std::vector<double> dy(g.nyp, 0.0);
    for (int iy = 1; iy <= ny - 1; ++iy) dy[iy + 1] = 0.5 * (y[iy + 2] - y[iy]);  // dnsdata.f90:155
    CHB_CUDA_OK(cudaMemcpy(h->t_y, y, sizeof(double) * g.nyp, cudaMemcpyHostToDevice));
    CHB_CUDA_OK(cudaMemcpy(h->t_dy, dy.data(), sizeof(double) * g.nyp, cudaMemcpyHostToDevice));
    const double* src[4] = {d0, d1, d2, d4};
    double* dst[4] = {h->t_d0, h->t_d1, h->t_d2, h->t_d4};
    for (int t = 0; t < 4; ++t) {
        std::vector<double> full((size_t)g.nyp * 5, 0.0);
        memcpy(&full[(size_t)2 * 5], src[t], sizeof(double) * 5 * (ny - 1));  // rows iy=1..ny-1 -> index iy+1
        CHB_CUDA_OK(cudaMemcpy(dst[t], full.data(), sizeof(double) * full.size(), cudaMemcpyHostToDevice));
    }
    CHB_CUDA_OK(cudaMemcpy(h->t_D0mat, D0mat, sizeof(double) * 5 * (ny + 1), cudaMemcpyHostToDevice));
    DevTables& t = h->tab;
    t.y = h->t_y; t.dy = h->t_dy; t.d0 = h->t_d0; t.d1 = h->t_d1; t.d2 = h->t_d2; t.d4 = h->t_d4;
    t.D0mat = h->t_D0mat;
    t.rows = h->t_rows;
#define CP5(name) memcpy(t.name, name, sizeof(double) * 5)
    CP5(d140); CP5(d14m1); CP5(d240); CP5(d24m1); CP5(d14n); CP5(d14np1); CP5(d24n); CP5(d24np1);
    CP5(v0bc); CP5(v0m1bc); CP5(vnbc); CP5(vnp1bc); CP5(eta0bc); CP5(eta0m1bc); CP5(etanbc); CP5(etanp1bc);
#undef CP5
    h->tables_set = true;
    return 0;
}

// ---- field transfer --------------------------------------------------------------------------
// Host <-> device transfer of a 3-component field.  Fortran layout: the field crosses PCIe in x-slabs that are staged in
// the work arena (dead between the sweeps of buildrhs) and transposed on the device; everything is enqueued on the
// handle's stream in order and waited for once.
static int transfer_V(chb_handle h, double* host, bool upload, bool fortran_layout, cplx* field) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    const Geometry& g = h->g;
    const size_t fld = (size_t)g.nyp * g.M;
    if (!fortran_layout) {
        const cudaMemcpyKind kind = upload ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        if (upload) CHB_CUDA_OK(cudaMemcpyAsync(field, host, 3 * fld * sizeof(cplx), kind, h->stream));
        else CHB_CUDA_OK(cudaMemcpyAsync(host, field, 3 * fld * sizeof(cplx), kind, h->stream));
    } else {
        cplx* stage = reinterpret_cast<cplx*>(h->arena + h->stage_off);
        const size_t cap = (h->arena_bytes - h->stage_off) / sizeof(cplx);
        const size_t per_ix = (size_t)g.nzt * g.nyp;
        CHB_REQUIRE(cap >= per_ix, "transfer: work arena smaller than one x-mode of the field (raise CHB_WORK_GB)");
        const int halves = cap >= 2 * per_ix ? 2 : 1;
        int nix_max = (int)((cap / halves) / per_ix);
        if (nix_max > g.nxB) nix_max = g.nxB;
        int k = 0;
        for (int c = 0; c < 3; ++c) {
            for (int ix0 = 0; ix0 < g.nxB; ix0 += nix_max, ++k) {
                const int nix = (ix0 + nix_max <= g.nxB) ? nix_max : g.nxB - ix0;
                cplx* st = stage + (size_t)(k % halves) * (cap / halves);
                cplx* hc = reinterpret_cast<cplx*>(host) + ((size_t)c * g.nxB + ix0) * per_ix;
                const size_t bytes = (size_t)nix * per_ix * sizeof(cplx);
                if (upload) {
                    CHB_CUDA_OK(cudaMemcpyAsync(st, hc, bytes, cudaMemcpyHostToDevice, h->stream));
                    launch_fortran_to_planes(h, st, field + c * fld, ix0, nix, h->stream);
                } else {
                    launch_planes_to_fortran(h, field + c * fld, st, ix0, nix, h->stream);
                    CHB_CUDA_OK(cudaMemcpyAsync(hc, st, bytes, cudaMemcpyDeviceToHost, h->stream));
                }
            }
        }
    }
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int chb_host_register(void* ptr, size_t bytes) {
    CHB_REQUIRE(ptr && bytes, "chb_host_register: null argument");
    CHB_CUDA_OK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return 0;
}
extern "C" int chb_host_unregister(void* ptr) {
    CHB_REQUIRE(ptr, "chb_host_unregister: null argument");
    CHB_CUDA_OK(cudaHostUnregister(ptr));
    return 0;
}
extern "C" int chb_upload_V(chb_handle h, const double* V) { return transfer_V(h, const_cast<double*>(V), true, true, h ? h->V : nullptr); }
extern "C" int chb_download_V(chb_handle h, double* V) { return transfer_V(h, V, false, true, h ? h->V : nullptr); }
extern "C" int chb_upload_V_planes(chb_handle h, const double* V) { return transfer_V(h, const_cast<double*>(V), true, false, h ? h->V : nullptr); }
extern "C" int chb_download_V_planes(chb_handle h, double* V) { return transfer_V(h, V, false, false, h ? h->V : nullptr); }
// Generic body-force path (chb_set_body_force_host): the caller evaluates its own set_body_force hook
// on the host (chb_download_V, its Fortran code) and hands the result over; chb_set_body_force is then a no-op and
// buildrhs uses F as it is (ghost extension dnsdata.f90:616-629 included).  Host layout = Fortran F(-1:ny+1,-nz:nz,nx0:nxN,1:3).
extern "C" int chb_upload_F(chb_handle h, const double* F) {
    CHB_REQUIRE(h && F, "chb_upload_F: null argument");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (!h->F && dev_alloc(&h->F, (size_t)3 * h->g.nyp * h->g.M)) return 1;
    h->bf.enabled = 2;   // external: chb_set_body_force does not recompute F
    return transfer_V(h, const_cast<double*>(F), true, true, h->F);
}
extern "C" int chb_download_F(chb_handle h, double* F) {
    CHB_REQUIRE(h && h->F, "chb_download_F: body force not enabled");
    return transfer_V(h, F, false, true, h->F);
}
extern "C" int chb_download_F_planes(chb_handle h, double* F) {
    CHB_REQUIRE(h && h->F, "chb_download_F_planes: body force not enabled");
    return transfer_V(h, F, false, false, h->F);
}

static int download_n(chb_handle h, double* host, const cplx* dev, int ncomp) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    const size_t n = (size_t)ncomp * h->g.nyp * h->g.M;
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaMemcpy(host, dev, n * sizeof(cplx), cudaMemcpyDeviceToHost));
    return 0;
}
// Between chb_buildrhs and chb_linsolve the RHS of the eta / D2v equations sits in V components 0 / 1 (rows 1..ny-1).
extern "C" int chb_download_rhs(chb_handle h, double* p) { return download_n(h, p, h ? h->V : nullptr, 2); }
// The spectral products exist one chunk of planes at a time; chb_debug_capture_products(h, 1) makes the following
// sweeps keep a copy of every chunk (6 complex per point of extra device memory: tests and diagnostics only).
extern "C" int chb_debug_capture_products(chb_handle h, int on) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (on && !h->P_dbg) {
        if (dev_alloc(&h->P_dbg, (size_t)6 * h->g.nyp * h->g.M)) return 1;
        h->dev_bytes += (size_t)6 * h->g.nyp * h->g.M * sizeof(cplx);
    } else if (!on && h->P_dbg) {
        CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
        cudaFree(h->P_dbg);
        h->P_dbg = nullptr;
        h->dev_bytes -= (size_t)6 * h->g.nyp * h->g.M * sizeof(cplx);
    }
    return 0;
}
extern "C" int chb_download_products(chb_handle h, double* p) {
    CHB_REQUIRE(h && h->P_dbg, "chb_download_products: call chb_debug_capture_products(h, 1) before chb_buildrhs");
    return download_n(h, p, h->P_dbg, 6);
}

// ---- scalars ---------------------------------------------------------------------------------
static int push_scalars(chb_handle h) {
    CHB_CUDA_OK(cudaMemcpyAsync(h->sc, h->sc_host, sizeof(DevScalars), cudaMemcpyHostToDevice, h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}
static int pull_scalars(chb_handle h) {
    CHB_CUDA_OK(cudaMemcpyAsync(h->sc_host, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int chb_set_wall_velocity(chb_handle h, double u0, double uN) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (pull_scalars(h)) return 1;
    h->sc_host->u0 = u0;
    h->sc_host->uN = uN;
    return push_scalars(h);
}

extern "C" int chb_set_forcing(chb_handle h, double meanpx, double meanpz, double meanflowx, double meanflowz, int CPI,
                               int CPI_type, double gamma) {
    CHB_REQUIRE(h, "null handle");
    CHB_REQUIRE(!CPI || CPI_type == 0 || CPI_type == 1, "Wrong selection of CPI_Type");  // linsolve_blocking.inc:93-95
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (pull_scalars(h)) return 1;
    DevScalars* s = h->sc_host;
    s->meanpx = meanpx; s->meanpz = meanpz; s->meanflowx = meanflowx; s->meanflowz = meanflowz;
    s->CPI = CPI; s->CPI_type = CPI_type; s->gamma = gamma;
    return push_scalars(h);
}

// ---- body force ------------------------------------------------------------------------------
static int set_body_force_common(chb_handle h, int enable, const double* A, int exclude_mean) {
    CHB_CUDA_OK(cudaSetDevice(h->device));
    h->bf.enabled = enable;
    if (!enable) return 0;
    memcpy(h->bf.A, A, sizeof(double) * 9);
    h->bf.exclude_mean = exclude_mean;
    if (!h->F) {
        if (dev_alloc(&h->F, (size_t)3 * h->g.nyp * h->g.M)) return 1;
    } else {
        // a hook only assigns inside its mask: values a previous hook (or chb_upload_F) left outside the new mask must
        // not survive the reconfiguration, the reference's F starts from 0 (dnsdata.f90:146)
        CHB_CUDA_OK(cudaMemsetAsync(h->F, 0, (size_t)3 * h->g.nyp * h->g.M * sizeof(cplx), h->stream));
    }
    return 0;
}
extern "C" int chb_set_body_force_linear(chb_handle h, int enable, const double* A, const double* mask_y,
                                         const double* mask_z, int exclude_mean) {
    CHB_REQUIRE(h, "null handle");
    CHB_REQUIRE(!enable || (A && mask_y && mask_z), "chb_set_body_force_linear: null argument");
    if (set_body_force_common(h, enable, A, exclude_mean)) return 1;
    if (!enable) return 0;
    const Geometry& g = h->g;
    if (!h->bf.mask_y && (dev_alloc(&h->bf.mask_y, (size_t)g.nyp) || dev_alloc(&h->bf.mask_z, (size_t)g.nzt))) return 1;
    CHB_CUDA_OK(cudaMemcpy(h->bf.mask_y, mask_y, sizeof(double) * g.nyp, cudaMemcpyHostToDevice));
    CHB_CUDA_OK(cudaMemcpy(h->bf.mask_z, mask_z, sizeof(double) * g.nzt, cudaMemcpyHostToDevice));
    if (h->bf.mask_yz) {   // back to the separable form
        cudaFree(h->bf.mask_yz);
        h->bf.mask_yz = nullptr;
    }
    return 0;
}
extern "C" int chb_set_body_force_linear_yz(chb_handle h, int enable, const double* A, const double* mask_yz,
                                            int exclude_mean) {
    CHB_REQUIRE(h, "null handle");
    CHB_REQUIRE(!enable || (A && mask_yz), "chb_set_body_force_linear_yz: null argument");
    if (set_body_force_common(h, enable, A, exclude_mean)) return 1;
    if (!enable) return 0;
    const Geometry& g = h->g;
    if (!h->bf.mask_yz && dev_alloc(&h->bf.mask_yz, (size_t)g.nyp * g.nzt)) return 1;
    CHB_CUDA_OK(cudaMemcpy(h->bf.mask_yz, mask_yz, sizeof(double) * g.nyp * g.nzt, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int chb_set_body_force(chb_handle h) {
    CHB_REQUIRE(h, "null handle");
    if (!h->bf.enabled || h->bf.enabled == 2) return 0;   // 2: F comes from the host (chb_upload_F)
    CHB_CUDA_OK(cudaSetDevice(h->device));
    launch_body_force(h);
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- the hot path ----------------------------------------------------------------------------
// One sweep of `convolutions` over all planes (dnsdata.f90:487-602), chunk by chunk, and - when rhs_ode is given - the
// plane loop of buildrhs (:634-670) behind it.  Chunk c uses lane c % nlanes.
//
// One lane (one GPU, two GPUs, NCCL fallback): everything on the handle's stream,
//     zfwd(c) -> barrier -> [convvel] -> xpass(c) -> barrier -> zbwd(c) -> rhs(c)          for c = 0, 1, ...
//
// Two lanes (default from 4 GPUs on): a software pipeline over two streams that sit on disjoint SM partitions when
// green contexts are available (green_ctx.cu):
//     sA:  [convvel(c)] xpass(c) -> barrier                      the FP64-heavy x-pass, whose stores are the xTOz transpose
//     sB:  zfwd(c+1) -> barrier -> zbwd(c) -> rhs(c)             the HBM-bound z-passes (zfwd's stores are zTOx) and the RHS
// so that the x-pass of chunk c runs beside the z-passes of its neighbours: NVLink carries zTOx(c+1) and xTOz(c) at
// the same time, HBM serves zbwd / rhs while the x-pass computes, and the SMs that one kernel leaves idle while its
// remote stores drain work for the other stream.  This is the role of the reference's nonblockingXZ variant
// (mpi_transpose.f90:149-168: MPI_IAlltoall progressing under the next plane's FFTs).  Measured at 4 GPUs
// (profiles/r2_multi_gpu.md): the first version, transposes (zfwd + xpass) on one partition and local kernels on the
// other, lost to the sequential sweep because zfwd leaves its SMs idle and the x-pass is compute-bound on a partition.
//
// Buffer reuse across ranks: a peer's zfwd(c) stores into this rank's Ar(lane) - free once every rank has passed the
// barrier behind xpass(c - nlanes) (sB waits for evA of the lane); a peer's xpass(c) stores into Br(lane) - it starts
// behind the barrier after zfwd(c), at which this rank only arrives once its own zbwd(c - nlanes) has read Br(lane)
// (earlier on sB).
static int convolutions_all(chb_handle h, int compute_cfl, bool products, const double* rhs_ode = nullptr, double rhs_deltat = 0.0,
                            double deltat = 0.0) {
    // the first sweep of buildrhs after an outstats also feeds the convection-velocity diagnostic (dnsdata.f90:515-531)
    const bool convvel = products && h->cv_enabled && h->cv_compute;
    const Geometry& g = h->g;
    const int np = h->chunk_planes;
    const int nch = (g.nyp + np - 1) / np;
    auto first = [&](int c) { return c * np; };
    auto count = [&](int c) { return (c + 1) * np <= g.nyp ? np : g.nyp - c * np; };
    auto local_part = [&](int c, cudaStream_t st) -> int {   // zbwd(c) -> rhs(c) on st, lane of chunk c selected
        Lane& ln = h->lane[c % h->nlanes];
        h->cstream = st;
        launch_zbwd(h, first(c), count(c));
        if (h->P_dbg)   // debug capture: keep this chunk's products
            for (int k = 0; k < 6; ++k)
                CHB_CUDA_OK(cudaMemcpyAsync(h->P_dbg + ((size_t)k * g.nyp + first(c)) * g.M, ln.Pc + (size_t)k * np * g.M,
                                            (size_t)count(c) * g.M * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
        if (rhs_ode) launch_rhs_chunk(h, rhs_ode, rhs_deltat, first(c), count(c), st);
        return 0;
    };
    if (h->sA == h->stream) {   // one lane, one stream
        h->cstream = h->stream;
        for (int c = 0; c < nch; ++c) {
            chb_select_lane(h, c % h->nlanes);
            launch_zfwd(h, first(c), count(c));          // stores straight into the x-side owner's buffer
            if (chb_exchange(h, true)) return 1;         // zTOx, mpi_transpose.f90:50-83
            if (convvel) launch_convvel(h, first(c), count(c), deltat);
            launch_xpass(h, first(c), count(c), compute_cfl);
            // xTOz, mpi_transpose.f90:88-117; in direct mode the barrier also frees Ar for the next chunk
            if ((products || h->p2p) && chb_exchange(h, false)) return 1;
            if (products && local_part(c, h->stream)) return 1;
        }
    } else {
        // fork: both streams start after everything queued on the main stream (V complete)
        CHB_CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
        CHB_CUDA_OK(cudaStreamWaitEvent(h->sA, h->ev_fork, 0));
        CHB_CUDA_OK(cudaStreamWaitEvent(h->sB, h->ev_fork, 0));
        auto zfwd_part = [&](int c) -> int {   // sB: zfwd(c) -> barrier; evZ of the lane
            Lane& ln = h->lane[c % h->nlanes];
            chb_select_lane(h, c % h->nlanes);
            h->cstream = h->sB;
            if (c >= h->nlanes) CHB_CUDA_OK(cudaStreamWaitEvent(h->sB, ln.evA, 0));   // Ar(lane) read by every rank's xpass(c - nlanes)
            launch_zfwd(h, first(c), count(c));
            if (chb_exchange(h, true)) return 1;
            CHB_CUDA_OK(cudaEventRecord(ln.evZ, h->sB));
            return 0;
        };
        if (zfwd_part(0)) return 1;
        for (int c = 0; c < nch; ++c) {
            Lane& ln = h->lane[c % h->nlanes];
            chb_select_lane(h, c % h->nlanes);
            h->cstream = h->sA;
            CHB_CUDA_OK(cudaStreamWaitEvent(h->sA, ln.evZ, 0));
            if (convvel) launch_convvel(h, first(c), count(c), deltat);
            launch_xpass(h, first(c), count(c), compute_cfl);
            if (chb_exchange(h, false)) return 1;
            CHB_CUDA_OK(cudaEventRecord(ln.evA, h->sA));
            if (c + 1 < nch && zfwd_part(c + 1)) return 1;
            if (products) {
                chb_select_lane(h, c % h->nlanes);
                CHB_CUDA_OK(cudaStreamWaitEvent(h->sB, ln.evA, 0));
                if (local_part(c, h->sB)) return 1;
            }
        }
        // join
        CHB_CUDA_OK(cudaEventRecord(h->ev_join, h->sA));
        CHB_CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        CHB_CUDA_OK(cudaEventRecord(h->ev_join, h->sB));
        CHB_CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    if (convvel) {   // IF (iy==nyN+2 .AND. compute_convvel): convvel_cnt=convvel_cnt+1; compute_convvel=.FALSE.   :546-549
        h->cv_cnt += 1;
        h->cv_compute = 0;
    }
    h->cstream = h->sA;
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int chb_cfl_prepass(chb_handle h) {
    CHB_REQUIRE(h && h->tables_set, "chb_cfl_prepass: tables not set");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (convolutions_all(h, 1, false)) return 1;
    launch_meanflow_prepass(h);
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int chb_buildrhs(chb_handle h, const double* ode, double deltat, int compute_cfl) {
    CHB_REQUIRE(h && h->tables_set, "chb_buildrhs: tables not set");
    CHB_REQUIRE(deltat > 0.0, "chb_buildrhs: deltat must be > 0");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (h->bf.enabled) launch_force_ghosts(h);
    if (convolutions_all(h, compute_cfl, true, ode, deltat, deltat)) return 1;
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int chb_linsolve(chb_handle h, double lambda) {
    CHB_REQUIRE(h && h->tables_set, "chb_linsolve: tables not set");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    launch_linsolve(h, lambda);
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int chb_vetaTOuvw(chb_handle h) { CHB_REQUIRE(h, "null handle"); return 0; }
extern "C" int chb_computeflowrate(chb_handle h, double) { CHB_REQUIRE(h, "null handle"); return 0; }

extern "C" int chb_rk3_step(chb_handle h, double deltat) {
    static const double RK[3][3] = {{120.0 / 32.0, 2.0, 0.0},                    // dnsdata.f90:70-72
                                    {120.0 / 8.0, 50.0 / 8.0, 34.0 / 8.0},
                                    {120.0 / 20.0, 90.0 / 20.0, 50.0 / 20.0}};
    for (int k = 0; k < 3; ++k) {
        if (chb_set_body_force(h)) return 1;
        if (chb_buildrhs(h, RK[k], deltat, k == 2)) return 1;
        if (chb_linsolve(h, RK[k][0] / deltat)) return 1;
    }
    return 0;
}

extern "C" int chb_get_step_scalars(chb_handle h, double* cfl, double* fr, double* corrpx, double* corrpz,
                                    double* meanpx, double* meanpz, double* U_lo, double* U_hi, double* W_lo,
                                    double* W_hi) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (h->g.nranks > 1) {
        if (chb_allreduce_max_cfl(h)) return 1;   // MPI_Allreduce(cfl, MAX), dnsdata.f90:861
        if (chb_bcast_scalars(h)) return 1;
    }
    if (pull_scalars(h)) return 1;
    if (chb_p2p_check(h)) return 6;
    chb_timer_flush(h);
    if (h->cv_enabled) h->cv_compute = 1;   // compute_convvel=.TRUE.   dnsdata.f90:858-860
    DevScalars* s = h->sc_host;
    double c;
    memcpy(&c, &s->cfl_bits, sizeof(double));
    if (cfl) *cfl = c;
    if (fr) memcpy(fr, s->fr, sizeof(double) * 3);
    if (corrpx) *corrpx = s->corrpx;
    if (corrpz) *corrpz = s->corrpz;
    if (meanpx) *meanpx = s->meanpx;
    if (meanpz) *meanpz = s->meanpz;
    if (U_lo) memcpy(U_lo, s->U_lo, sizeof(double) * 5);
    if (U_hi) memcpy(U_hi, s->U_hi, sizeof(double) * 5);
    if (W_lo) memcpy(W_lo, s->W_lo, sizeof(double) * 5);
    if (W_hi) memcpy(W_hi, s->W_hi, sizeof(double) * 5);
    // cfl = 0 (dnsdata.f90:861)
    CHB_CUDA_OK(cudaMemsetAsync(&h->sc->cfl_bits, 0, sizeof(unsigned long long), h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- diagnostics -----------------------------------------------------------------------------
extern "C" int chb_sync(chb_handle h) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaGetLastError());
    if (chb_p2p_check(h)) return 6;
    return 0;
}
extern "C" int chb_stopwatch_begin(chb_handle h) {
    CHB_REQUIRE(h, "null handle");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (!h->sw0) {
        CHB_CUDA_OK(cudaEventCreate(&h->sw0));
        CHB_CUDA_OK(cudaEventCreate(&h->sw1));
    }
    CHB_CUDA_OK(cudaEventRecord(h->sw0, h->stream));
    return 0;
}
extern "C" int chb_stopwatch_end(chb_handle h, double* ms) {
    CHB_REQUIRE(h && h->sw0 && ms, "chb_stopwatch_end: stopwatch not started");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    CHB_CUDA_OK(cudaEventRecord(h->sw1, h->stream));
    CHB_CUDA_OK(cudaEventSynchronize(h->sw1));
    float f = 0;
    CHB_CUDA_OK(cudaEventElapsedTime(&f, h->sw0, h->sw1));
    *ms = f;
    CHB_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int chb_get_stream(chb_handle h, void** stream) {
    CHB_REQUIRE(h && stream, "null argument");
    *stream = (void*)h->stream;
    return 0;
}
extern "C" long long chb_device_bytes(chb_handle h) { return h ? (long long)h->dev_bytes : -1; }
extern "C" long long chb_launch_count(chb_handle h) { return h ? h->launches : -1; }
extern "C" int chb_timing_enable(chb_handle h, int on) {
    CHB_REQUIRE(h, "null handle");
    chb_timer_flush(h);
    h->timer.on = on != 0;
    h->timer.recs.clear();
    return 0;
}
extern "C" int chb_timing_report(chb_handle h, char* names, int name_stride, double* ms, long long* launches, int cap) {
    CHB_REQUIRE(h, "null handle");
    chb_timer_flush(h);
    int i = 0;
    for (auto& kv : h->timer.recs) {
        if (i < cap) {
            snprintf(names + (size_t)i * name_stride, name_stride, "%s", kv.first.c_str());
            ms[i] = kv.second.ms;
            launches[i] = kv.second.n;
        }
        ++i;
    }
    return i;
}

// standalone FFT for the parity tests (conv_kernels.cu)
int launch_test_fft(const FftPlan& pl, const cplx* W, const int* rev, cplx* data, int nlines, int sign);
extern "C" int chb_test_fft_lines(int n, int nlines, int sign, double* data_host) {
    FftPlan pl;
    std::vector<int> rev;
    CHB_REQUIRE(build_plan(n, &pl, &rev), "chb_test_fft_lines: n must be 2^a or 3*2^a");
    int ndev = 0;
    CHB_CUDA_OK(cudaGetDeviceCount(&ndev));
    CHB_REQUIRE(ndev > 0, "no CUDA device");
    std::vector<double> w = twiddles(n, n, n);
    double* dW = nullptr;
    int* dRev = nullptr;
    cplx* dData = nullptr;
    if (dev_upload(&dW, w) || dev_upload(&dRev, rev)) return 1;
    CHB_CUDA_OK(cudaMalloc((void**)&dData, sizeof(cplx) * (size_t)n * nlines));
    CHB_CUDA_OK(cudaMemcpy(dData, data_host, sizeof(cplx) * (size_t)n * nlines, cudaMemcpyHostToDevice));
    if (launch_test_fft(pl, (const cplx*)dW, dRev, dData, nlines, sign)) return 1;
    CHB_CUDA_OK(cudaDeviceSynchronize());
    CHB_CUDA_OK(cudaMemcpy(data_host, dData, sizeof(cplx) * (size_t)n * nlines, cudaMemcpyDeviceToHost));
    cudaFree(dW); cudaFree(dRev); cudaFree(dData);
    return 0;
}
