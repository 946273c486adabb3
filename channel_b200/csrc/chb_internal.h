// chb_internal.h - handle layout and kernel launchers of libchannel_b200.
#pragma once
#define CHB_SOLVE_K 8
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
#include "fft_device.cuh"
#include "transpose_index.h"

// Coefficient tables of setup_derivatives / setup_boundary_conditions
// (dnsdata.f90:241-308), passed by value to the y-direction kernels.
struct DevTables {
    const double* y;      // [ny+3]      y(-1:ny+1)
    const double* dy;     // [ny+3]      dy(iy), iy=1..ny-1 (else 0)
    const double* d0;     // [ny+3][5]   der(iy)%d0(-2:2), zero rows outside 1..ny-1
    const double* d1;
    const double* d2;
    const double* d4;
    const double* D0mat;  // [ny+1][5]   after LU5decompStep, row i <-> iy=i+1
    const double* rows;   // [ny+3][5][5] k2-polynomial coefficients of the D2vmat / etamat rows of this substep (solve_device.cuh)
    double d140[5], d14m1[5], d240[5], d24m1[5], d14n[5], d14np1[5], d24n[5], d24np1[5];
    double v0bc[5], v0m1bc[5], vnbc[5], vnp1bc[5], eta0bc[5], eta0m1bc[5], etanbc[5], etanp1bc[5];
};

// Device-resident scalars of MODULE dnsdata that the path reads and writes.
struct DevScalars {
    unsigned long long cfl_bits;  // running max of cfl (non-negative double, ordered as uint64)
    double fr[3];
    double corrpx, corrpz;
    double meanpx, meanpz;
    double meanflowx, meanflowz;
    double gamma;
    double u0, uN;
    int CPI, CPI_type;
    double U_lo[5], U_hi[5], W_lo[5], W_hi[5];
};

struct Geometry {
    int nx, ny, nz, nxd, nzd;
    int nyp;        // ny+3 planes (iy=-1..ny+1)
    int nzt;        // 2nz+1
    int rank, nranks;
    int nx0, nxN, nxB;
    int nz0, nzN, nzB;
    long long M;    // local columns = nxB*nzt
    double alfa0, beta0, ni;
    double dx, dz, factor;
    int tw;         // log2 of the x-tile width of the products work buffer (transpose_index.h)
    int twa;        // same for the velocity work buffer; < 0 = row-major
};

struct BodyForce {
    int enabled;
    double A[9];
    double* mask_y;  // [ny+3]
    double* mask_z;  // [2nz+1]
    double* mask_yz; // [ny+3][2nz+1] or null: general mask(iy,iz), used instead of mask_y*mask_z
    int exclude_mean;
};

// Where the pack side of a pencil transpose stores: one base pointer per destination rank.
// Direct NVLink mode: the peer's receive buffer, mapped through CUDA IPC (the kernels write the
// transposed data straight into peer HBM, no pack buffer and no separate all-to-all).  NCCL mode:
// this rank's send buffer.  Single GPU: the local receive buffer.  In all modes the element for
// peer q lives at p[q] + index(block = this rank, ...).
#define CHB_MAX_RANKS 8
struct PeerPtrs {
    cplx* p[CHB_MAX_RANKS];
};

// One set of pencil-transpose work buffers for a chunk of y-planes.  Consecutive chunks alternate between the lanes
// (chb_api.cu, convolutions_all): while the x-pass works on chunk c in one lane, the z-passes and the plane loop of
// buildrhs work on chunks c+1 / c-1 in the other.  All lanes are carved out of one device allocation, the arena (one CUDA IPC handle per rank).
#define CHB_MAX_LANES 2
struct Lane {
    cplx *A, *Ar, *B, *Br;   // A, B: NCCL mode only (send buffers)
    cplx* Pc;                // spectral products of the chunk [6][np][M]: written by zbwd, read by the plane loop of buildrhs
    PeerPtrs Aw, Bw;
    unsigned long long* flags;
    unsigned long long* peer_flags[CHB_MAX_RANKS];
    unsigned long long epoch;
    cudaEvent_t evZ;         // recorded on sB after zfwd + its barrier: Ar complete on every rank
    cudaEvent_t evA;         // recorded on sA after xpass + its barrier: Br complete, Ar free on every rank
};

struct KernelTimer {
    bool on = false;
    struct Rec { double ms = 0; long long n = 0; };
    std::map<std::string, Rec> recs;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
};

struct chb_handle_s {
    Geometry g;
    int device;
    cudaStream_t stream;
    cudaStream_t side_stream;      // mean-mode column, concurrent with S3/S4 of the solve
    cudaEvent_t ev_fork, ev_join;
    Lane lane[CHB_MAX_LANES];      // the fields A..epoch below are a copy of the lane in use (chb_select_lane)
    int nlanes, cur_lane;          // CHB_LANES = 1 | 2 (default 2 on several GPUs)
    cudaStream_t sA, sB;           // x-pass / z-passes + RHS assembly of the chunk pipeline; both = stream when nlanes == 1
    cudaStream_t cstream;          // stream the conv launchers use (set by convolutions_all to sA or sB)
    void* green[2];                // CUgreenCtx of sA / sB when the SMs are partitioned (CHB_GREEN, green_ctx.cu), else null
    int green_sms[2];              // SMs of each partition (0 = not partitioned)
    // fields (device layout [c][iy+1][ixl][iz+nz], complex128)
    cplx* V;        // [3][nyp][M]; between chb_buildrhs and chb_linsolve components 0 / 1 hold the RHS of the eta / D2v
                    // equations on rows 1..ny-1 (the reference's in-place write-back, dnsdata.f90:667-671)
    cplx* oldrhs;   // [2][nyp][M]
    cplx* F;        // [3][nyp][M] or null
    double* ckpt;   // [nblk][8][M]  UL-recurrence state every CHB_SOLVE_K rows (solve_kernels.cu)
    // work arena: barrier flags + per lane Ar, Br, Pc (+ A, B in NCCL mode); also the staging area of the host <-> device
    // field transfers and of the restart files
    char* arena;
    size_t arena_bytes;
    size_t stage_off;     // first byte of the arena usable as staging (behind the flags and the Ar buffers, which peers may write)
    int chunk_planes;
    int zf_lines_per_cta, zb_lines_per_cta;  // lines per CTA of zfwd / zbwd (CHB_ZF_LPC, CHB_ZB_LPC: 2, 4 or 8)
    int use_fft3;         // register-resident three-stage FFT kernels for the large sizes (CHB_FFT3=0 disables)
    int zf_direct;        // CHB_ZF_DIRECT=1: zfwd4 stage A reads global memory directly (no TMA staging); measured slower
    int z_tpl;            // CHB_Z_TPL=128|96: threads per line of the z passes at nzd = 1536 / 3072 (default 64)
    int solve_pf;         // CHB_SOLVE_PF=1: S1 / S3 / S4 with eight rows of loads in flight per thread; measured slower at config 3
    double* rhs_state;    // [32][M] accumulators of the plane loop of buildrhs carried from chunk to chunk
    int xpass_split;      // CHB_XPASS_SPLIT: two threads per innermost butterfly position (default on at nxd = 1536)
    int xpass_persist;    // CHB_XPASS_PERSIST: persistent x-pass with the next line's inputs prefetched into shared memory
    cplx* A;        // NCCL mode only: send buffer of zTOx, [peer][3][np][nzB][nxB]
    cplx* Ar;       // z-padded velocity after zTOx, [src rank][3][np][nzB][nxB]
    cplx* B;        // NCCL mode only: send buffer of xTOz
    cplx* Br;       // products after xTOz, [src rank][6][np][nxB/2^tw][nzB][2^tw]
    cplx* Pc;       // spectral products of the chunk in flight (lane)
    PeerPtrs Aw, Bw;   // where zfwd / xpass store (see PeerPtrs)
    int p2p;           // 1 = direct NVLink stores into peer buffers + flag barrier, 0 = NCCL all-to-all
    unsigned long long* flags;             // [CHB_MAX_RANKS] barrier flags of this rank (IPC-shared)
    unsigned long long* peer_flags[CHB_MAX_RANKS];
    unsigned long long* p2p_error;         // device word set by a barrier that timed out
    void* ipc_opened[CHB_MAX_RANKS];
    int n_ipc_opened;
    // debug capture of the spectral products of all planes (chb_debug_capture_products): [6][nyp][M] or null
    cplx* P_dbg;
    // FFT plans and tables
    FftPlan plan_z, plan_x;
    cplx* Wz;       // exp(+2 pi i e/nzd)
    cplx* Wx;       // exp(+2 pi i e/nxd)
    cplx* Wh;       // exp(+ i pi k/nxd), k=0..nxd/2
    int* rev_z;     // digit reversal for plan_z
    // tables
    DevTables tab;
    double *t_y, *t_dy, *t_d0, *t_d1, *t_d2, *t_d4, *t_D0mat, *t_rows;
    bool tables_set;
    DevScalars* sc;       // device
    DevScalars* sc_host;  // pinned
    BodyForce bf;
    // multi-GPU
    void* nccl_comm;
    // convection-velocity diagnostic (convvel.cu; #ifdef convvel of the reference)
    int cv_enabled, cv_compute;   // compute: set by chb_get_step_scalars (outstats), consumed by the next buildrhs sweep
    long long cv_cnt;             // convvel_cnt
    cplx* cv_Vold;                // Voldz  [3][ny+3][nzd][nxB]
    double* cv_uconv;             // uconv  [3][ny+3][nzd][nxB]
    // restart / snapshot files (restart_io.cu)
    void* rio;
    double grid_a, grid_ymin, grid_ymax;   // dns.in a, ymin, ymax: header fields of Dati.cart.out
    // bookkeeping
    long long launches;
    size_t dev_bytes;
    cudaEvent_t sw0, sw1;
    KernelTimer timer;
};

// doubles of dynamic shared memory of mean_mode_kernel (solve_kernels.cu): A [ny+1][5] | ucor, U, W [ny+3] | wts [ny/2+1][3]
inline size_t mean_mode_smem_doubles(int ny) { return (size_t)(ny + 1) * 5 + 3 * (size_t)(ny + 3) + 3 * (size_t)(ny / 2 + 1); }

// ---- launchers (conv_kernels.cu) ----
void launch_zfwd(chb_handle_s* h, int plane0, int nplanes);
void launch_xpass(chb_handle_s* h, int plane0, int nplanes, int compute_cfl);
void launch_zbwd(chb_handle_s* h, int plane0, int nplanes);
// ---- zpass3_kernels.cu / xpass3_kernels.cu: false = no specialised kernel for this size ----
bool launch_z3_fwd_or_bwd(chb_handle_s* h, int plane0, int nplanes, bool fwd);
bool launch_x3_pass(chb_handle_s* h, int plane0, int nplanes, int compute_cfl);
// ---- rhs_kernel.cu ----
void launch_rhs_chunk(chb_handle_s* h, const double* ode, double deltat, int plane0, int nplanes, cudaStream_t st);
// ---- solve_kernels.cu ----
void launch_linsolve(chb_handle_s* h, double lambda);
void launch_meanflow_prepass(chb_handle_s* h);
// ---- layout_kernels.cu ----
// x-slab [ix0, ix0 + nix) of one component between the file / Fortran order (cols: [nix][2nz+1][ny+3], contiguous) and the
// device layout (planes: the component's [ny+3][nxB][2nz+1] array), on stream st
void launch_fortran_to_planes(chb_handle_s* h, const cplx* cols, cplx* planes, int ix0, int nix, cudaStream_t st);
void launch_planes_to_fortran(chb_handle_s* h, const cplx* planes, cplx* cols, int ix0, int nix, cudaStream_t st);
void launch_body_force(chb_handle_s* h);
void launch_force_ghosts(chb_handle_s* h);
// ---- transposes (transpose.cu) ----
int chb_alltoall(chb_handle_s* h, const cplx* send, cplx* recv, size_t count_per_peer);
int chb_p2p_setup(chb_handle_s* h);                         // maps the peers' arenas (Ar/Br/flags of every lane); 0 on success
int chb_p2p_check(chb_handle_s* h);                         // non-zero (and the error text) if a flag barrier timed out
int chb_allreduce_min_i64(chb_handle_s* h, long long* v, int n);   // agreement of the ranks on sizes derived from local state
// ---- green_ctx.cu: SM partitions for the two streams of the chunk pipeline (CUDA green contexts) ----
int chb_green_create(chb_handle_s* h, int sms_a);           // 0 = sA / sB now run on disjoint SM partitions
void chb_green_destroy(chb_handle_s* h);
void chb_p2p_teardown(chb_handle_s* h);
int chb_exchange(chb_handle_s* h, bool a_side);             // completes zTOx (a_side) / xTOz on the lane's stream
void chb_select_lane(chb_handle_s* h, int lane);            // makes `lane` the one the conv launchers use

// ---- convvel.cu ----
void launch_convvel(chb_handle_s* h, int plane0, int nplanes, double deltat);
// ---- restart_io.cu ----
void chb_restart_destroy(chb_handle_s* h);

// Kernel launch.  The device build expands to the <<< >>> syntax; tests/host_emul builds the same sources with g++
// and routes launches to the CTA emulator (test infrastructure, never part of the product library).  The kernel
// goes last so that template arguments with commas need no parentheses:
//   CHB_LAUNCH(grid, block, smem_bytes, stream, kernel<args>)(kernel arguments);
#ifdef CHB_HOST_EMUL
#define CHB_LAUNCH(grid, block, smem, stream, ...) cta_emul::launcher(__VA_ARGS__, grid, block)
#else
#define CHB_LAUNCH(grid, block, smem, stream, ...) __VA_ARGS__<<<grid, block, smem, stream>>>
#endif

// timing helpers
struct ScopedKernelTimer {
    chb_handle_s* h;
    const char* name;
    cudaEvent_t e0, e1;
    bool on;
    cudaStream_t st;
    ScopedKernelTimer(chb_handle_s* h_, const char* name_, cudaStream_t st_ = nullptr);
    ~ScopedKernelTimer();
};
void chb_timer_flush(chb_handle_s* h);

#define CHB_CUDA_OK(call)                                                         \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            chb_set_error(std::string(#call) + ": " + cudaGetErrorString(e__));   \
            return 1;                                                             \
        }                                                                         \
    } while (0)
void chb_set_error(const std::string& s);
