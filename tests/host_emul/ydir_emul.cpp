// ydir_emul.cpp - test infrastructure: the y-direction CUDA kernels (rhs_kernel.cu, solve_kernels.cu,
// bodyforce_kernels.cu) compiled with g++ and run on CPU threads (cta_emul.hpp).
//
// Those kernels are one thread per wavenumber column with no shared memory, shuffles or atomics, so
// their source is valid host C++ once the CUDA keywords are neutralised; running it here lets the
// CPU test suite (-m "not gpu") check the kernel logic (rhs_kernel, solve_s1..s4, mean_mode_kernel)
// against the numpy oracle without a GPU.  It is a checker only: the product path never loads this library.
#include "cta_emul.hpp"

#include <cmath>
#include <cstring>
#include <vector>

#define CHB_HOST_EMUL 1
#include "../../channel_b200/csrc/rhs_kernel.cu"
#include "../../channel_b200/csrc/solve_kernels.cu"
#include "../../channel_b200/csrc/bodyforce_kernels.cu"

// run kern(args...) on a 1-D grid of 1-D blocks (cta_emul.hpp: one OS thread per CUDA thread, real barriers)
template <class K, class... Args>
static void emulate(K kern, int blocks, int threads, Args... args) {
    cta_emul::launch(kern, dim3(blocks, 1, 1), threads, args...);
}
template <class K, class... Args>
static void emulate2(K kern, int gx, int gy, int threads, Args... args) {
    cta_emul::launch(kern, dim3(gx, gy, 1), threads, args...);
}

extern "C" {

// One substep of the y-direction work on host arrays in the device layout [c][iy+1][ixl][iz+nz]:
//   V [3][nyp][M] (in: u,v,w; out: u,v,w after linsolve), P [6][nyp][M] products, F [3][nyp][M] or null,
//   oldrhs [2][nyp][M] (in/out), rhs_out [2][nyp][M] (plain flow: the RHS; fused flow: Step1 results).
// Tables as chb_set_tables receives them.  scal_io: {meanpx, meanpz, meanflowx, meanflowz, gamma, u0, uN,
// CPI, CPI_type} in; {fr0, fr1, fr2, corrpx, corrpz, meanpx} and, from index 10, U_lo, U_hi, W_lo, W_hi (5 each) out; 30 doubles.  mode: 0 whole substep, -1 rhs_kernel only, 2 / -2 the same with the chunked rhs march, 3 the whole substep with the prefetching S1 / S3 / S4.
__attribute__((visibility("default"))) int chb_emul_ydir_substep(int nx, int ny, int nz, double alfa0, double beta0, double ni, const double* y,
                          const double* d0, const double* d1, const double* d2, const double* d4,
                          const double* bc5x16, const double* D0mat, double* V, const double* P, const double* F,
                          double* oldrhs, double* rhs_out, double* scal_io, double ode1, double ode2, double ode3,
                          double deltat, int mode) {
    Geometry g;
    memset(&g, 0, sizeof(g));
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.nyp = ny + 3; g.nzt = 2 * nz + 1;
    g.rank = 0; g.nranks = 1;
    g.nx0 = 0; g.nxN = nx; g.nxB = nx + 1;
    g.M = (long long)g.nxB * g.nzt;
    g.alfa0 = alfa0; g.beta0 = beta0; g.ni = ni;
    const int nyp = g.nyp;
    // tables, as chb_set_tables lays them out (chb_api.cu)
    std::vector<double> dy(nyp, 0.0), full[4];
    for (int iy = 1; iy <= ny - 1; ++iy) dy[iy + 1] = 0.5 * (y[iy + 2] - y[iy]);
    const double* src[4] = {d0, d1, d2, d4};
    for (int t = 0; t < 4; ++t) {
        full[t].assign((size_t)nyp * 5, 0.0);
        memcpy(&full[t][10], src[t], sizeof(double) * 5 * (ny - 1));
    }
    DevTables tab;
    tab.y = y; tab.dy = dy.data();
    tab.d0 = full[0].data(); tab.d1 = full[1].data(); tab.d2 = full[2].data(); tab.d4 = full[3].data();
    tab.D0mat = D0mat;
    std::vector<double> rows((size_t)nyp * 25, 0.0);
    tab.rows = rows.data();
    double* dst[16] = {tab.d140, tab.d14m1, tab.d240, tab.d24m1, tab.d14n, tab.d14np1, tab.d24n, tab.d24np1,
                       tab.v0bc, tab.v0m1bc, tab.vnbc, tab.vnp1bc, tab.eta0bc, tab.eta0m1bc, tab.etanbc, tab.etanp1bc};
    for (int k = 0; k < 16; ++k) memcpy(dst[k], bc5x16 + 5 * k, sizeof(double) * 5);
    DevScalars sc;
    memset(&sc, 0, sizeof(sc));
    sc.meanpx = scal_io[0]; sc.meanpz = scal_io[1]; sc.meanflowx = scal_io[2]; sc.meanflowz = scal_io[3];
    sc.gamma = scal_io[4]; sc.u0 = scal_io[5]; sc.uN = scal_io[6];
    sc.CPI = (int)scal_io[7]; sc.CPI_type = (int)scal_io[8];

    const size_t fld = (size_t)nyp * g.M;
    std::vector<double> ckpt((size_t)((ny - 1) / CHB_SOLVE_K + 1) * 8 * g.M, 0.0);
    cplx* Vc = reinterpret_cast<cplx*>(V);
    const cplx* Pc = reinterpret_cast<const cplx*>(P);
    const cplx* Fc = reinterpret_cast<const cplx*>(F);
    cplx* oc = reinterpret_cast<cplx*>(oldrhs);
    cplx* rc = reinterpret_cast<cplx*>(rhs_out);
    const int T = 128, blocks = (int)((g.M + T - 1) / T);
    const double lam = ode1 / deltat;
    {
        // the plane loop of buildrhs works in place: finished rows go to V components 0 / 1 (dnsdata.f90:667-671)
        std::vector<double> state((size_t)32 * g.M, 0.0);
        if (mode == 2 || mode == -2) {   // chunked march with carried accumulators: uneven chunks of input planes, each
                                         // with its own products array [6][n][M] as the library's lanes hold them
            const int cuts[5] = {-1, 2, 3, ny / 2, ny + 2};    // chunks [-1,1], [2,2], [3,ny/2-1], [ny/2, ny+1]
            for (int c = 0; c < 4; ++c) {
                const int ip0 = cuts[c], ip1 = cuts[c + 1] - 1, n = ip1 - ip0 + 1, plane0 = ip0 + 1;
                std::vector<cplx> Pk((size_t)6 * n * g.M);
                for (int k = 0; k < 6; ++k)
                    memcpy(&Pk[(size_t)k * n * g.M], Pc + ((size_t)k * nyp + plane0) * g.M, sizeof(cplx) * (size_t)n * g.M);
                if (F) emulate(rhs_kernel<true, 3>, blocks, T, Vc, (const cplx*)Pk.data(), Fc, oc, g, tab, &sc, lam, ode2, ode3, ip0, ip1, state.data(), n, plane0);
                else emulate(rhs_kernel<false, 3>, blocks, T, Vc, (const cplx*)Pk.data(), Fc, oc, g, tab, &sc, lam, ode2, ode3, ip0, ip1, state.data(), n, plane0);
            }
        } else {
            if (F) emulate(rhs_kernel<true, 3>, blocks, T, Vc, Pc, Fc, oc, g, tab, &sc, lam, ode2, ode3, -1, ny + 1, state.data(), nyp, 0);
            else emulate(rhs_kernel<false, 3>, blocks, T, Vc, Pc, Fc, oc, g, tab, &sc, lam, ode2, ode3, -1, ny + 1, state.data(), nyp, 0);
        }
        memcpy(rc, Vc, sizeof(cplx) * 2 * fld);   // what chb_download_rhs returns
        if (mode < 0) return 0;   // RHS only
        emulate(solve_rows_kernel, (nyp * 5 + 127) / 128, 128, tab, rows.data(), lam, ni, nyp);
        if (mode == 3) {   // S1 / S3 / S4 with eight rows of loads in flight per thread
            emulate(solve_s1_kernel<0, 8>, blocks, T, Vc, ckpt.data(), g, tab, &sc, lam);
            emulate(solve_s1_kernel<1, 8>, blocks, T, Vc, ckpt.data(), g, tab, &sc, lam);
        } else {
            emulate(solve_s1_kernel<0, 1>, blocks, T, Vc, ckpt.data(), g, tab, &sc, lam);
            emulate(solve_s1_kernel<1, 1>, blocks, T, Vc, ckpt.data(), g, tab, &sc, lam);
        }
        emulate(solve_s2_kernel<0>, blocks, T, (const double*)ckpt.data(), Vc, g, tab, &sc, lam);
        emulate(solve_s2_kernel<1>, blocks, T, (const double*)ckpt.data(), Vc, g, tab, &sc, lam);
        if (mean_mode_smem_doubles(ny) * sizeof(double) > sizeof(cta_emul::g_dyn_smem)) return 3;
        emulate(mean_mode_kernel, 1, MEAN_THREADS, Vc, g, tab, &sc, lam);
        if (mode == 3) {
            emulate(solve_s3_kernel<8>, blocks, T, Vc, g, tab);
            emulate(solve_s4_kernel<8>, blocks, T, Vc, g, tab);
        } else {
            emulate(solve_s3_kernel<1>, blocks, T, Vc, g, tab);
            emulate(solve_s4_kernel<1>, blocks, T, Vc, g, tab);
        }
    }
    scal_io[0] = sc.fr[0]; scal_io[1] = sc.fr[1]; scal_io[2] = sc.fr[2];
    scal_io[3] = sc.corrpx; scal_io[4] = sc.corrpz; scal_io[5] = sc.meanpx;
    for (int i = 0; i < 5; ++i) {   // the wall values outstats reads (dnsdata.f90:866-870)
        scal_io[10 + i] = sc.U_lo[i]; scal_io[15 + i] = sc.U_hi[i]; scal_io[20 + i] = sc.W_lo[i]; scal_io[25 + i] = sc.W_hi[i];
    }
    return 0;
}

// set_body_force (body_force_kernel) followed, if ghosts != 0, by the ghost-node extension of buildrhs
// (force_ghost_kernel, dnsdata.f90:616-629).  V, F: [3][nyp][M] device layout; masks as chb_set_body_force_linear
// (mask_yz null) or chb_set_body_force_linear_yz (mask_yz set); d4: [(ny-1)][5] as chb_set_tables receives it.
__attribute__((visibility("default"))) int chb_emul_body_force(int nx, int ny, int nz, const double* V, double* F,
                                                               const double* A, const double* mask_y, const double* mask_z,
                                                               const double* mask_yz, int exclude_mean, int ghosts,
                                                               const double* d4) {
    Geometry g;
    memset(&g, 0, sizeof(g));
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.nyp = ny + 3; g.nzt = 2 * nz + 1;
    g.nranks = 1; g.nxN = nx; g.nxB = nx + 1;
    g.M = (long long)g.nxB * g.nzt;
    BodyForce bf;
    memset(&bf, 0, sizeof(bf));
    bf.enabled = 1;
    memcpy(bf.A, A, sizeof(double) * 9);
    bf.mask_y = const_cast<double*>(mask_y);
    bf.mask_z = const_cast<double*>(mask_z);
    bf.mask_yz = const_cast<double*>(mask_yz);
    bf.exclude_mean = exclude_mean;
    emulate2(body_force_kernel, (int)((g.M + 255) / 256), g.nyp, 256, reinterpret_cast<const cplx*>(V),
             reinterpret_cast<cplx*>(F), g, bf);
    if (ghosts) {
        std::vector<double> full((size_t)g.nyp * 5, 0.0);
        memcpy(&full[10], d4, sizeof(double) * 5 * (ny - 1));
        DevTables tab;
        memset(&tab, 0, sizeof(tab));
        tab.d4 = full.data();
        emulate(force_ghost_kernel, (int)((g.M + 255) / 256), 256, reinterpret_cast<cplx*>(F), g, tab);
    }
    return 0;
}

}  // extern "C"
