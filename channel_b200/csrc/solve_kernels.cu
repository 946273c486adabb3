// solve_kernels.cu - per-(ix,iz) pentadiagonal solves for v and eta and the recovery of u,w
// (linsolve, linsolve_blocking.inc:3-107; LU5decompStep / LeftLU5divStep1 / LeftLU5divStep2,
// rbparmat_blocking.f90:20-100 with npy=1; COMPLEXderiv, dnsdata.f90:339-373).
//
// One thread per column; the wavenumber index is the contiguous one, so every load/store of a
// warp is a 512-byte coalesced segment.  The right-hand sides arrive in V components 0 (eta) and 1 (v), rows
// 1..ny-1, where buildrhs left them (dnsdata.f90:667-671), and every sweep works in place on V.  Four sweeps:
//   S1 (iy = ny-1 -> 1): build the D2vmat / etamat rows (linsolve_blocking.inc:12-13), fold the
//       wall BCs (applybc_0/n, dnsdata.f90:458-472), UL-factorise on the fly and apply
//       LeftLU5divStep1 to both right-hand sides; the two L-multipliers per row and matrix go to
//       `mult`, the intermediate x overwrites the rhs array.
//   S2 (iy = 1 -> ny-1): LeftLU5divStep2, then the ghost-node closures (:50-61).
//   S3 (iy = ny+1 -> -1): compact first derivative of v: d1 stencil, one-sided wall rows, wall
//       corrections and LeftLU5divStep1 with the pre-factorised D0mat (COMPLEXderiv).
//   S4 (iy = -1 -> ny+1): LeftLU5divStep2 with D0mat, then u=(ia*vy-ib*eta)/k2, w=(ib*vy+ia*eta)/k2
//       (linsolve_blocking.inc:99-103).
// The mean column (0,0) is finished by a single-thread kernel (linsolve_blocking.inc:62-97).
#include <cstdlib>
#include "chb_internal.h"
#include "solve_device.cuh"

#define SOLVE_THREADS 128

// k2-polynomial coefficients of the matrix rows for this substep's lam (solve_device.cuh)
__global__ void solve_rows_kernel(DevTables tab, double* __restrict__ rows, double lam, double ni, int nyp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (iy+1)*5 + j
    if (i >= nyp * 5) return;
    const double d0 = tab.d0[i], d2 = tab.d2[i], d4 = tab.d4[i];
    double* t = rows + (size_t)i * 5;
    t[0] = lam * d2 - ni * d4;
    t[1] = 2.0 * ni * d2 - lam * d0;
    t[2] = -ni * d0;
    t[3] = lam * d0 - ni * d2;
    t[4] = ni * d0;
}

// ---------------------------------------------------------------------------------------------
// COMP = 0: eta equation (etamat), COMP = 1: v equation (D2vmat); blockIdx.y selects nothing, the
// two components are separate launches of the same grid so that each thread carries one recurrence
// (half the registers, twice the resident warps).
// PF (CHB_SOLVE_PF=1; measured slower at 524 k and at 262 k columns per GPU, kept as a comparator): the loads of the next PF rows are issued before the rows are processed (as S2
// does with its blocks), so that a thread keeps PF instead of one or two 16-byte loads in flight: for the
// strong-scaled runs, where a GPU has too few columns to hide the latency with threads alone.
// CG: the in-place row traffic with ld.global.cg / st.global.cg (L2 only), as in S2
template <int COMP, int PF = 1, bool CG = false>
__global__ void __launch_bounds__(SOLVE_THREADS)
solve_s1_kernel(cplx* __restrict__ rhs, double* __restrict__ ckpt, Geometry g, DevTables tab,
                const DevScalars* __restrict__ sc, double lam) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    const int ix = g.nx0 + ixl, iz = izp - g.nz;
    const double al = g.alfa0 * ix, be = g.beta0 * iz;
    const double k2 = al * al + be * be;
    const bool mean = (ix == 0 && iz == 0);
    const int ny = g.ny;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    // wall data: only the mean mode carries non-zero BCs (bc%eta = u0 / uN), linsolve_blocking.inc:17,31
    double bc0_eta = 0.0, bcn_eta = 0.0;
    if (COMP == 0 && mean) {
        bc0_eta = sc->u0;
        bcn_eta = sc->uN;
    }
    const double* bcn = COMP ? tab.vnbc : tab.etanbc;
    const double* bcnp1 = COMP ? tab.vnp1bc : tab.etanp1bc;
    const double* bc0 = COMP ? tab.v0bc : tab.eta0bc;
    const double* bc0m1 = COMP ? tab.v0m1bc : tab.eta0m1bc;
    cplx* __restrict__ col = rhs + COMP * comp + m;
    LUState st = {0, 0, 0, 0};
    cplx x1 = make_double2(0, 0), x2 = x1;  // x(i+1), x(i+2)
    auto row = [&](int iy, cplx b) {
        Row5 r;
        build_row_poly<COMP>(tab.rows, iy, k2, r);
        const size_t off = (size_t)(iy + 1) * plane;
        if (iy == ny - 1) {
            fold_top1(r, bcn, bcnp1);
            if (COMP == 0) b.x -= r.a[3] * bcn_eta / tab.etanbc[3];  // linsolve_blocking.inc:40
            r.a[3] = r.a[4] = 0.0;                                   // rbparmat_blocking.f90:29
        } else if (iy == ny - 2) {
            fold_top2(r, bcn);
            if (COMP == 0) b.x -= r.a[4] * bcn_eta / tab.etanbc[3];  // :41
            r.a[4] = 0.0;
        }
        if (iy == 1) {
            fold_bot1(r, bc0, bc0m1);
            if (COMP == 0) b.x -= r.a[1] * bc0_eta / tab.eta0bc[1];  // :26
        } else if (iy == 2) {
            fold_bot2(r, bc0);
            if (COMP == 0) b.x -= r.a[0] * bc0_eta / tab.eta0bc[1];  // :27
        }
        double inv, u1, u2;
        lu_row(r, st, inv, u1, u2);
        // LeftLU5divStep1 (rbparmat_blocking.f90:70-72)
        cplx x;
        x.x = (b.x - (u1 * x1.x + u2 * x2.x)) * inv;
        x.y = (b.y - (u1 * x1.y + u2 * x2.y)) * inv;
        x2 = x1;
        x1 = x;
        if constexpr (CG) __stcg(reinterpret_cast<double2*>(col + off), x); else col[off] = x;
        // The L-multipliers Step2 needs are not stored: the state of the UL recurrence is
        // checkpointed every SOLVE_K rows and solve_s2_kernel recomputes them block by block.
        if (iy > 1 && (iy - 1) % SOLVE_K == 0) {
            double* ck = ckpt + ((size_t)((iy - 1) / SOLVE_K - 1) * 8 + (COMP ? 0 : 4)) * plane + m;
            ck[0 * plane] = st.l1m2; ck[1 * plane] = st.l1m1; ck[2 * plane] = st.l2m2; ck[3 * plane] = st.l2m1;
        }
    };
    if constexpr (PF > 1) {
        for (int i0 = ny - 1; i0 >= 1; i0 -= PF) {
            cplx bb[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) bb[k] = (i0 - k >= 1) ? col[(size_t)(i0 - k + 1) * plane] : make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < PF; ++k)
                if (i0 - k >= 1) row(i0 - k, bb[k]);
        }
    } else {
        for (int iy = ny - 1; iy >= 1; --iy) {
            cplx b;
            if constexpr (CG) b = __ldcg(reinterpret_cast<const double2*>(col + (size_t)(iy + 1) * plane)); else b = col[(size_t)(iy + 1) * plane];
            row(iy, b);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// MEMV: cache operators of the in-place row traffic (0: ld.global.nc / st; 1: ld.cg / st; 2: ld.cg / st.cg; 3: ld.cs / st.cs)
template <int MEMV>
__device__ __forceinline__ cplx s2_load(const cplx* p) {
    if constexpr (MEMV == 0) return __ldg(p);
    else if constexpr (MEMV == 3) { const double2 v = __ldcs(reinterpret_cast<const double2*>(p)); return v; }
    else { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return v; }
}
template <int MEMV>
__device__ __forceinline__ void s2_store(cplx* p, cplx v) {
    if constexpr (MEMV == 2) __stcg(reinterpret_cast<double2*>(p), v);
    else if constexpr (MEMV == 3) __stcs(reinterpret_cast<double2*>(p), v);
    else *p = v;
}

template <int COMP, int MEMV = 0>
__global__ void __launch_bounds__(SOLVE_THREADS, 4)
solve_s2_kernel(const double* __restrict__ ckpt, cplx* V, Geometry g,
                DevTables tab, const DevScalars* __restrict__ sc, double lam) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    const int ix = g.nx0 + ixl, iz = izp - g.nz;
    const double al = g.alfa0 * ix, be = g.beta0 * iz;
    const double k2 = al * al + be * be;
    const bool mean = (ix == 0 && iz == 0);
    const int ny = g.ny;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    double bc0_w = 0.0, bcn_w = 0.0;   // wall value of the unknown: bc%eta for the mean mode, else 0
    if (COMP == 0 && mean) {
        bc0_w = sc->u0;
        bcn_w = sc->uN;
    }
    const double* bcn = COMP ? tab.vnbc : tab.etanbc;
    const double* bcnp1 = COMP ? tab.vnp1bc : tab.etanp1bc;
    const double* bc0 = COMP ? tab.v0bc : tab.eta0bc;
    const double* bc0m1 = COMP ? tab.v0m1bc : tab.eta0m1bc;
    // in place: the Step1 result of a block of rows is read (prefetched) before the block's rows are written
    const cplx* xin = V + COMP * comp + m;
    cplx* out = V + COMP * comp + m;
    cplx v1 = make_double2(0, 0), v2 = v1, v3 = v1;  // b(i-1), b(i-2), b(i-3)
    for (int i0 = 1; i0 <= ny - 1; i0 += SOLVE_K) {
        // ---- recompute the multipliers of rows i0..i0+K-1 (descending, as solve_s1 did) ----
        LUState st = {0, 0, 0, 0};
        if (i0 + SOLVE_K <= ny - 1) {   // state after row i0+K was processed
            const double* ck = ckpt + ((size_t)((i0 - 1) / SOLVE_K) * 8 + (COMP ? 0 : 4)) * plane + m;
            st.l1m2 = ck[0 * plane]; st.l1m1 = ck[1 * plane]; st.l2m2 = ck[2 * plane]; st.l2m1 = ck[3 * plane];
        }
        // prefetch the Step1 results of this block while the multipliers are being recomputed
        cplx xb[SOLVE_K];
#pragma unroll
        for (int k = 0; k < SOLVE_K; ++k) {
            const int iy = i0 + k;
            // read-only path: a row is read before this thread (the only one that touches the column) overwrites it, and
            // the compiler may hoist the next block's loads over this block's stores as it could with separate arrays
            xb[k] = (iy <= ny - 1) ? s2_load<MEMV>(xin + (size_t)(iy + 1) * plane) : make_double2(0.0, 0.0);
        }
        double m2[SOLVE_K], m1[SOLVE_K];
#pragma unroll
        for (int k = SOLVE_K - 1; k >= 0; --k) {
            const int iy = i0 + k;
            m2[k] = m1[k] = 0.0;
            if (iy <= ny - 1) {
                Row5 r;
                build_row_poly<COMP>(tab.rows, iy, k2, r);
                if (iy == ny - 1) {
                    fold_top1(r, bcn, bcnp1);
                    r.a[3] = r.a[4] = 0.0;
                } else if (iy == ny - 2) {
                    fold_top2(r, bcn);
                    r.a[4] = 0.0;
                }
                if (iy == 1) fold_bot1(r, bc0, bc0m1);
                else if (iy == 2) fold_bot2(r, bc0);
                double inv, u1, u2;
                lu_row(r, st, inv, u1, u2);
                // rbparmat_blocking.f90:45 zeroes A(1,-2:-1), A(2,-2)
                if (iy >= 3) m2[k] = st.l1m2;
                if (iy >= 2) m1[k] = st.l1m1;
            }
        }
        // ---- LeftLU5divStep2 (rbparmat_blocking.f90:93-95), ascending ----
#pragma unroll
        for (int k = 0; k < SOLVE_K; ++k) {
            const int iy = i0 + k;
            if (iy <= ny - 1) {
                cplx v = xb[k];
                v.x -= m2[k] * v2.x + m1[k] * v1.x;
                v.y -= m2[k] * v2.y + m1[k] * v1.y;
                s2_store<MEMV>(out + (size_t)(iy + 1) * plane, v);
                v3 = v2; v2 = v1; v1 = v;
                if (iy == 3) {  // bottom closure needs nodes 1..3 (linsolve_blocking.inc:51-54)
                    const cplx a1 = v3, a2 = v2, a3 = v1;
                    cplx vw, vg;
                    vw.x = (bc0_w - (a1.x * bc0[2] + a2.x * bc0[3] + a3.x * bc0[4])) / bc0[1];
                    vw.y = (0.0 - (a1.y * bc0[2] + a2.y * bc0[3] + a3.y * bc0[4])) / bc0[1];
                    if (COMP) {
                        vg.x = (0.0 - (vw.x * bc0m1[1] + a1.x * bc0m1[2] + a2.x * bc0m1[3] + a3.x * bc0m1[4])) / bc0m1[0];
                        vg.y = (0.0 - (vw.y * bc0m1[1] + a1.y * bc0m1[2] + a2.y * bc0m1[3] + a3.y * bc0m1[4])) / bc0m1[0];
                    } else {
                        vg.x = -(vw.x * bc0m1[1] + a1.x * bc0m1[2] + a2.x * bc0m1[3] + a3.x * bc0m1[4]) / bc0m1[0];
                        vg.y = -(vw.y * bc0m1[1] + a1.y * bc0m1[2] + a2.y * bc0m1[3] + a3.y * bc0m1[4]) / bc0m1[0];
                    }
                    out[1 * plane] = vw;
                    out[0 * plane] = vg;
                }
            }
        }
    }
    {  // top closure (linsolve_blocking.inc:57-60): nodes ny-3..ny-1 = v3,v2,v1
        cplx vw, vg;
        vw.x = (bcn_w - (v3.x * bcn[0] + v2.x * bcn[1] + v1.x * bcn[2])) / bcn[3];
        vw.y = (0.0 - (v3.y * bcn[0] + v2.y * bcn[1] + v1.y * bcn[2])) / bcn[3];
        if (COMP) {
            vg.x = (0.0 - (v3.x * bcnp1[0] + v2.x * bcnp1[1] + v1.x * bcnp1[2] + vw.x * bcnp1[3])) / bcnp1[4];
            vg.y = (0.0 - (v3.y * bcnp1[0] + v2.y * bcnp1[1] + v1.y * bcnp1[2] + vw.y * bcnp1[3])) / bcnp1[4];
        } else {
            vg.x = -(v3.x * bcnp1[0] + v2.x * bcnp1[1] + v1.x * bcnp1[2] + vw.x * bcnp1[3]) / bcnp1[4];
            vg.y = -(v3.y * bcnp1[0] + v2.y * bcnp1[1] + v1.y * bcnp1[2] + vw.y * bcnp1[3]) / bcnp1[4];
        }
        out[(size_t)(ny + 1) * plane] = vw;
        out[(size_t)(ny + 2) * plane] = vg;
    }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx stencil5(const double* c, cplx a0, cplx a1, cplx a2, cplx a3, cplx a4) {
    cplx r;
    r.x = c[0] * a0.x + c[1] * a1.x + c[2] * a2.x + c[3] * a3.x + c[4] * a4.x;
    r.y = c[0] * a0.y + c[1] * a1.y + c[2] * a2.y + c[3] * a3.y + c[4] * a4.y;
    return r;
}

// S3: vy (before Step2) -> V comp 3
template <int PF = 1>
__global__ void __launch_bounds__(SOLVE_THREADS)
solve_s3_kernel(cplx* __restrict__ V, Geometry g, DevTables tab) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    if (g.nx0 + ixl == 0 && izp == g.nz) return;  // mean column: no vetaTOuvw
    const int ny = g.ny;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    const cplx* v = V + 1 * comp + m;
    cplx* out = V + 2 * comp + m;
#define VAT(iy) v[(size_t)((iy) + 1) * plane]
    // top one-sided rows (dnsdata.f90:354-355)
    cplx w4 = VAT(ny + 1), w3 = VAT(ny), w2 = VAT(ny - 1), w1 = VAT(ny - 2), w0 = VAT(ny - 3);
    const cplx f_n = stencil5(tab.d14n, w0, w1, w2, w3, w4);
    const cplx f_np1 = stencil5(tab.d14np1, w0, w1, w2, w3, w4);
    out[(size_t)(ny + 1) * plane] = f_n;
    out[(size_t)(ny + 2) * plane] = f_np1;
    // bottom one-sided rows (dnsdata.f90:350-351)
    cplx f_0, f_m1;
    {
        const cplx b0 = VAT(-1), b1 = VAT(0), b2 = VAT(1), b3 = VAT(2), b4 = VAT(3);
        f_0 = stencil5(tab.d140, b0, b1, b2, b3, b4);
        f_m1 = stencil5(tab.d14m1, b0, b1, b2, b3, b4);
        out[(size_t)1 * plane] = f_0;
        out[(size_t)0 * plane] = f_m1;
    }
    cplx x1 = make_double2(0, 0), x2 = x1;
    // window w0..w4 = v(iy-2..iy+2); currently holds ny-3..ny+1 = window of iy = ny-1
    auto row = [&](int iy) {
        const int ti = (iy + 1) * 5;
        double c[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) c[j] = __ldg(&tab.d1[ti + j]);
        cplx f = stencil5(c, w0, w1, w2, w3, w4);  // dnsdata.f90:358
        if (iy == ny - 1) {                        // :365
            const double a = __ldg(&tab.d0[ti + 3]), b = __ldg(&tab.d0[ti + 4]);
            f.x -= a * f_n.x + b * f_np1.x;
            f.y -= a * f_n.y + b * f_np1.y;
        } else if (iy == ny - 2) {                 // :366
            const double b = __ldg(&tab.d0[ti + 4]);
            f.x -= b * f_n.x;
            f.y -= b * f_n.y;
        }
        if (iy == 1) {                             // :361
            const double a = __ldg(&tab.d0[ti + 1]), b = __ldg(&tab.d0[ti + 0]);
            f.x -= a * f_0.x + b * f_m1.x;
            f.y -= a * f_0.y + b * f_m1.y;
        } else if (iy == 2) {                      // :362
            const double b = __ldg(&tab.d0[ti + 0]);
            f.x -= b * f_0.x;
            f.y -= b * f_0.y;
        }
        // LeftLU5divStep1 with D0mat (row i = iy-1)
        const double* A = tab.D0mat + (size_t)(iy - 1) * 5;
        const double a0 = __ldg(&A[2]), a1 = __ldg(&A[3]), a2 = __ldg(&A[4]);
        cplx x;
        x.x = (f.x - (a1 * x1.x + a2 * x2.x)) * a0;
        x.y = (f.y - (a1 * x1.y + a2 * x2.y)) * a0;
        x2 = x1;
        x1 = x;
        out[(size_t)(iy + 1) * plane] = x;
    };
    if constexpr (PF > 1) {
        for (int i0 = ny - 1; i0 >= 1; i0 -= PF) {
            cplx nb[PF];   // the values that enter the window after rows i0, i0-1, ...
#pragma unroll
            for (int k = 0; k < PF; ++k) nb[k] = (i0 - k >= 1 && i0 - k - 3 >= -1) ? VAT(i0 - k - 3) : make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < PF; ++k)
                if (i0 - k >= 1) {
                    row(i0 - k);
                    w4 = w3; w3 = w2; w2 = w1; w1 = w0;
                    w0 = nb[k];
                }
        }
    } else {
        for (int iy = ny - 1; iy >= 1; --iy) {
            row(iy);
            // slide the window down
            w4 = w3; w3 = w2; w2 = w1; w1 = w0;
            if (iy - 3 >= -1) w0 = VAT(iy - 3);
        }
    }
#undef VAT
}

// S4: Step2 with D0mat, then u,w
template <int PF = 1, bool CG = false>
__global__ void __launch_bounds__(SOLVE_THREADS)
solve_s4_kernel(cplx* __restrict__ V, Geometry g, DevTables tab) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    const int ix = g.nx0 + ixl, iz = izp - g.nz;
    if (ix == 0 && iz == 0) return;
    const double al = g.alfa0 * ix, be = g.beta0 * iz;
    const double rk2 = 1.0 / (al * al + be * be);   // one division per column instead of four per node
    const int ny = g.ny;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    cplx b1 = make_double2(0, 0), b2 = b1;
    auto row = [&](int iy, cplx vy, const cplx eta) {
        const size_t off = (size_t)(iy + 1) * plane + m;
        if (iy >= 1) {  // rows i=0..ny <-> iy=1..ny+1 (rows ny, ny+1 of D0mat are zero)
            const double* A = tab.D0mat + (size_t)(iy - 1) * 5;
            const double am2 = __ldg(&A[0]), am1 = __ldg(&A[1]);
            vy.x -= am2 * b2.x + am1 * b1.x;
            vy.y -= am2 * b2.y + am1 * b1.y;
        }
        b2 = b1;
        b1 = vy;
        // (ia*vy - ib*eta)/k2 ; (ib*vy + ia*eta)/k2
        cplx u, w;
        u.x = (-al * vy.y + be * eta.y) * rk2;
        u.y = (al * vy.x - be * eta.x) * rk2;
        w.x = (-be * vy.y - al * eta.y) * rk2;
        w.y = (be * vy.x + al * eta.x) * rk2;
        if constexpr (CG) {
            __stcg(reinterpret_cast<double2*>(V + 0 * comp + off), u);
            __stcg(reinterpret_cast<double2*>(V + 2 * comp + off), w);
        } else {
            V[0 * comp + off] = u;
            V[2 * comp + off] = w;
        }
    };
    if constexpr (PF > 1) {
        for (int i0 = -1; i0 <= ny + 1; i0 += PF) {
            cplx vyb[PF], etab[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const bool in = i0 + k <= ny + 1;
                const size_t off = (size_t)(i0 + k + 1) * plane + m;
                vyb[k] = in ? V[2 * comp + off] : make_double2(0.0, 0.0);
                etab[k] = in ? V[0 * comp + off] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int k = 0; k < PF; ++k)
                if (i0 + k <= ny + 1) row(i0 + k, vyb[k], etab[k]);
        }
    } else {
        for (int iy = -1; iy <= ny + 1; ++iy) {
            const size_t off = (size_t)(iy + 1) * plane + m;
            if constexpr (CG) row(iy, __ldcg(reinterpret_cast<const double2*>(V + 2 * comp + off)), __ldcg(reinterpret_cast<const double2*>(V + 0 * comp + off)));
            else row(iy, V[2 * comp + off], V[0 * comp + off]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// yintegr (dnsdata.f90:312-324) on a real column f(iy), iy=-1..ny+1, element stride `st`
__device__ double yintegr_dev(const double* __restrict__ y, const double* f, size_t st, int ny, int imag_part) {
    double II = 0.0;
    for (int iy = 1; iy <= ny - 1; iy += 2) {
        const double yp1 = y[iy + 2] - y[iy + 1], ym1 = y[iy] - y[iy + 1];
        const double a1 = -1.0 / 3.0 * ym1 + 1.0 / 6.0 * yp1 + 1.0 / 6.0 * yp1 * yp1 / ym1;
        const double a3 = +1.0 / 3.0 * yp1 - 1.0 / 6.0 * ym1 - 1.0 / 6.0 * ym1 * ym1 / yp1;
        const double a2 = yp1 - ym1 - a1 - a3;
        II = II + a1 * f[(size_t)iy * st + imag_part] + a2 * f[(size_t)(iy + 1) * st + imag_part] +
             a3 * f[(size_t)(iy + 2) * st + imag_part];
    }
    return II;
}

__device__ void cpi_update(DevScalars* sc, double ni) {
    if (sc->CPI) {  // linsolve_blocking.inc:87-97, channel.f90:104-114
        if (sc->CPI_type == 0)
            sc->meanpx = (1.0 - sc->gamma) * 6.0 * ni / sc->fr[0];
        else if (sc->CPI_type == 1)
            sc->meanpx = (1.5 / sc->gamma) * sc->fr[0] * ni;
    }
}

__device__ void store_wall_columns(cplx* V, DevScalars* sc, const Geometry& g, size_t m00) {
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    for (int i = 0; i < 5; ++i) {
        sc->U_lo[i] = V[0 * comp + (size_t)i * plane + m00].x;
        sc->W_lo[i] = V[2 * comp + (size_t)i * plane + m00].x;
        sc->U_hi[i] = V[0 * comp + (size_t)(g.ny - 2 + i) * plane + m00].x;
        sc->W_hi[i] = V[2 * comp + (size_t)(g.ny - 2 + i) * plane + m00].x;
    }
}

// mean column (0,0) after S2: linsolve_blocking.inc:62-97.  One block, everything in shared memory.  Parallel over the
// threads: the strided accesses to the column (one 16-byte element per plane of M columns), the rows of etamat(0,0)
// with their wall-BC folding, the weights of yintegr.  Sequential, in the reference's order of operations, but as three
// independent chains on three warps: (a) UL factorisation of etamat(0,0) -> ucor solve -> its integral, (b) the integral
// of U, (c) the integral of W.  (As one thread reading tables and scratch from global memory this kernel took 2 ms per
// substep at ny = 512 - nothing on one GPU, where it hides under S3 / S4, but 7 % of the step on eight.)
// Dynamic shared memory (doubles): A [ny+1][5] | ucor [ny+3] | U [ny+3] | W [ny+3] | wts [3][ny/2+1].
#define MEAN_THREADS 128

// yintegr (dnsdata.f90:312-324) with the weights a1, a2, a3 of every odd node precomputed (same expressions)
__device__ __forceinline__ double yintegr_weighted(const double* __restrict__ wts, const double* f, int ny) {
    double II = 0.0;
    for (int iy = 1, k = 0; iy <= ny - 1; iy += 2, ++k)
        II = II + wts[3 * k + 0] * f[iy] + wts[3 * k + 1] * f[iy + 1] + wts[3 * k + 2] * f[iy + 2];
    return II;
}

__global__ void __launch_bounds__(MEAN_THREADS)
mean_mode_kernel(cplx* __restrict__ V, Geometry g, DevTables tab, DevScalars* sc, double lam) {
    if (blockIdx.x != 0) return;
    CHB_DYN_SMEM(double, sm);
    const int ny = g.ny, nz = g.nz, nyp = g.nyp;
    const int tid = threadIdx.x, nth = blockDim.x;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    const size_t m00 = (size_t)nz;  // ixl=0, izp=nz
    double* A = sm;                                  // [ny+1][5]: folded rows of etamat(0,0), then its UL factors
    double* ucor = A + (size_t)(ny + 1) * 5;         // [ny+3], index iy+1
    double* U = ucor + nyp;                          // Re eta -> u(0,0)
    double* W = U + nyp;                             // Im eta -> w(0,0)
    double* wts = W + nyp;                           // [ny/2+1][3]
    // V(:,0,0,3) = Im eta ; V(:,0,0,1) = Re eta            :63-64
    for (int i = tid; i < nyp; i += nth) {
        const cplx e = V[0 * comp + (size_t)i * plane + m00];
        V[2 * comp + (size_t)i * plane + m00] = make_double2(e.y, 0.0);
        V[0 * comp + (size_t)i * plane + m00] = make_double2(e.x, 0.0);
        U[i] = e.x;
        W[i] = e.y;
        ucor[i] = (i >= 2 && i <= ny) ? 1.0 : 0.0;   // rhs 1 on rows 1..ny-1                                :65-68
    }
    // rows of etamat(0,0) (k2 = 0) with the wall BCs folded in
    for (int iy = 1 + tid; iy <= ny - 1; iy += nth) {
        Row5 rv, re;
        build_rows(tab, iy, 0.0, lam, g.ni, rv, re);
        if (iy == ny - 1) { fold_top1(re, tab.etanbc, tab.etanp1bc); re.a[3] = re.a[4] = 0.0; }
        else if (iy == ny - 2) { fold_top2(re, tab.etanbc); re.a[4] = 0.0; }
        if (iy == 1) fold_bot1(re, tab.eta0bc, tab.eta0m1bc);
        else if (iy == 2) fold_bot2(re, tab.eta0bc);
        double* r = A + (size_t)(iy - 1) * 5;
        for (int j = 0; j < 5; ++j) r[j] = re.a[j];
    }
    for (int j = tid; j < 10; j += nth) A[(size_t)(ny - 1) * 5 + j] = 0.0;   // rows ny, ny+1 are zero
    // weights of yintegr at the odd nodes
    for (int k = tid; 1 + 2 * k <= ny - 1; k += nth) {
        const int iy = 1 + 2 * k;
        const double yp1 = tab.y[iy + 2] - tab.y[iy + 1], ym1 = tab.y[iy] - tab.y[iy + 1];
        const double a1 = -1.0 / 3.0 * ym1 + 1.0 / 6.0 * yp1 + 1.0 / 6.0 * yp1 * yp1 / ym1;
        const double a3 = +1.0 / 3.0 * yp1 - 1.0 / 6.0 * ym1 - 1.0 / 6.0 * ym1 * ym1 / yp1;
        wts[3 * k + 0] = a1;
        wts[3 * k + 1] = yp1 - ym1 - a1 - a3;
        wts[3 * k + 2] = a3;
    }
    __syncthreads();
    if (tid == 0) {
        // etamat(0,0), factorised again
        LUState st = {0, 0, 0, 0};
        for (int iy = ny - 1; iy >= 1; --iy) {
            double* r = A + (size_t)(iy - 1) * 5;
            Row5 re;
            for (int j = 0; j < 5; ++j) re.a[j] = r[j];
            double inv, u1, u2;
            lu_row(re, st, inv, u1, u2);
            r[0] = st.l1m2; r[1] = st.l1m1; r[2] = inv; r[3] = u1; r[4] = u2;
        }
        A[0] = A[1] = 0.0;  // rbparmat_blocking.f90:45
        A[5] = 0.0;
        for (int iy = ny - 1; iy >= 1; --iy) {
            const double* r = A + (size_t)(iy - 1) * 5;
            ucor[iy + 1] = (ucor[iy + 1] - (r[3] * ucor[iy + 2] + r[4] * ucor[iy + 3])) * r[2];
        }
        for (int iy = 1; iy <= ny + 1; ++iy) {
            const double* r = A + (size_t)(iy - 1) * 5;
            ucor[iy + 1] = ucor[iy + 1] - (r[0] * ucor[iy - 1] + r[1] * ucor[iy]);
        }
        {
            const double* e0bc = tab.eta0bc; const double* e0m1 = tab.eta0m1bc;
            const double* enbc = tab.etanbc; const double* enp1 = tab.etanp1bc;
            ucor[1] = -(ucor[2] * e0bc[2] + ucor[3] * e0bc[3] + ucor[4] * e0bc[4]) / e0bc[1];                      // :70
            ucor[0] = -(ucor[1] * e0m1[1] + ucor[2] * e0m1[2] + ucor[3] * e0m1[3] + ucor[4] * e0m1[4]) / e0m1[0];  // :71
            ucor[ny + 1] = -(ucor[ny - 2] * enbc[0] + ucor[ny - 1] * enbc[1] + ucor[ny] * enbc[2]) / enbc[3];      // :74
            ucor[ny + 2] = -(ucor[ny - 2] * enp1[0] + ucor[ny - 1] * enp1[1] + ucor[ny] * enp1[2] + ucor[ny + 1] * enp1[3]) / enp1[4];  // :75
        }
        sc->fr[2] = yintegr_weighted(wts, ucor, ny);
    } else if (tid == 32) {
        sc->fr[0] = yintegr_weighted(wts, U, ny);  // :77
    } else if (tid == 64) {
        sc->fr[1] = yintegr_weighted(wts, W, ny);
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
        if (fabs(sc->meanflowx) > 1.0e-7 && !sc->CPI) sc->corrpx = (sc->meanflowx - sc->fr[0]) / sc->fr[2];   // :79-82
        if (fabs(sc->meanflowz) > 1.0e-7 && !sc->CPI) sc->corrpz = (sc->meanflowz - sc->fr[1]) / sc->fr[2];   // :83-86
    }
    __threadfence_block();
    __syncthreads();
    if (fabs(sc->meanflowx) > 1.0e-7 && !sc->CPI) {
        const double c = sc->corrpx;
        for (int i = tid; i < nyp; i += nth) {
            U[i] += c * ucor[i];
            V[0 * comp + (size_t)i * plane + m00].x = U[i];
        }
    }
    if (fabs(sc->meanflowz) > 1.0e-7 && !sc->CPI) {
        const double c = sc->corrpz;
        for (int i = tid; i < nyp; i += nth) {
            W[i] += c * ucor[i];
            V[2 * comp + (size_t)i * plane + m00].x = W[i];
        }
    }
    __syncthreads();
    if (tid == 0) {
        cpi_update(sc, g.ni);
        for (int i = 0; i < 5; ++i) {   // what outstats reads (dnsdata.f90:866-870)
            sc->U_lo[i] = U[i];
            sc->W_lo[i] = W[i];
            sc->U_hi[i] = U[ny - 2 + i];
            sc->W_hi[i] = W[ny - 2 + i];
        }
    }
}

// channel.f90:101-115: flow rates of the initial mean profile and CPI meanpx
__global__ void meanflow_prepass_kernel(cplx* __restrict__ V, Geometry g, DevTables tab, DevScalars* sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    const size_t m00 = (size_t)g.nz;
    const double* Ucol = reinterpret_cast<const double*>(V + 0 * comp + m00);
    const double* Wcol = reinterpret_cast<const double*>(V + 2 * comp + m00);
    sc->fr[0] = yintegr_dev(tab.y, Ucol, 2 * plane, g.ny, 0);
    sc->fr[1] = yintegr_dev(tab.y, Wcol, 2 * plane, g.ny, 0);
    cpi_update(sc, g.ni);
    store_wall_columns(V, sc, g, m00);
}

#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)   // the kernel-only emulation harnesses (tests/host_emul) stop here
void launch_linsolve(chb_handle_s* h, double lam) {
    const Geometry& g = h->g;
    const int blocks = (int)((g.M + SOLVE_THREADS - 1) / SOLVE_THREADS);
    CHB_LAUNCH((g.nyp * 5 + 127) / 128, 128, 0, h->stream, solve_rows_kernel)(h->tab, h->t_rows, lam, g.ni, g.nyp);
    h->launches++;
    // the sweeps that work in place move their rows through L2 only (ld.global.cg / st.global.cg): reading rows through L1
    // that the same kernel overwrites cost S2 2 ms/step (19.4 -> 17.2), S1 / S4 0.4 ms (profiles/r2c_r2e_single_gpu.md)
    static const bool ycg = []() { const char* e = getenv("CHB_Y_CG"); return e ? atoi(e) != 0 : true; }();
    const bool pf = h->solve_pf != 0;   // eight rows of loads in flight per thread in S1 / S3 / S4 (measured slower, off)
    {
        ScopedKernelTimer tm(h, "solve_s1");
        if (pf) {
            CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s1_kernel<0, 8>)(h->V, h->ckpt, g, h->tab, h->sc, lam);
            CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s1_kernel<1, 8>)(h->V, h->ckpt, g, h->tab, h->sc, lam);
        } else {
            if (ycg) {
                CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s1_kernel<0, 1, true>)(h->V, h->ckpt, g, h->tab, h->sc, lam);
                CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s1_kernel<1, 1, true>)(h->V, h->ckpt, g, h->tab, h->sc, lam);
            } else {
                CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s1_kernel<0>)(h->V, h->ckpt, g, h->tab, h->sc, lam);
                CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s1_kernel<1>)(h->V, h->ckpt, g, h->tab, h->sc, lam);
            }
        }
    }
    {
        ScopedKernelTimer tm(h, "solve_s2");
        static const int memv = []() { const char* e = getenv("CHB_S2_MEMV"); return e ? atoi(e) : 2; }();   // measured: 19.4 / 17.5 / 17.2 / 17.2 ms for 0..3
#define CHB_S2(MV)                                                                                                            \
        CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s2_kernel<0, MV>)(h->ckpt, h->V, g, h->tab, h->sc, lam); \
        CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s2_kernel<1, MV>)(h->ckpt, h->V, g, h->tab, h->sc, lam)
        if (memv == 1) { CHB_S2(1); } else if (memv == 2) { CHB_S2(2); } else if (memv == 3) { CHB_S2(3); } else { CHB_S2(0); }
#undef CHB_S2
    }
    h->launches += 6;
    // The mean column (0,0) only needs the result of S2 and is skipped by S3/S4: finish it on the
    // side stream while S3/S4 run (it is a single-thread recurrence, linsolve_blocking.inc:62-97).
    const bool mean_here = (g.nx0 == 0);
    if (mean_here) {
        cudaEventRecord(h->ev_fork, h->stream);
        cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0);
        ScopedKernelTimer tm(h, "mean_mode", h->side_stream);
        const size_t msm = mean_mode_smem_doubles(g.ny) * sizeof(double);
        cudaFuncSetAttribute(mean_mode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
        CHB_LAUNCH(1, MEAN_THREADS, msm, h->side_stream, mean_mode_kernel)(h->V, g, h->tab, h->sc, lam);
        h->launches++;
    }
    {
        ScopedKernelTimer tm(h, "solve_s3");
        if (pf) CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s3_kernel<8>)(h->V, g, h->tab);
        else CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s3_kernel<1>)(h->V, g, h->tab);
    }
    {
        ScopedKernelTimer tm(h, "solve_s4");
        if (pf) CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s4_kernel<8>)(h->V, g, h->tab);
        else if (ycg) CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s4_kernel<1, true>)(h->V, g, h->tab);
        else CHB_LAUNCH(blocks, SOLVE_THREADS, 0, h->stream, solve_s4_kernel<1>)(h->V, g, h->tab);
    }
    if (mean_here) {
        cudaEventRecord(h->ev_join, h->side_stream);
        cudaStreamWaitEvent(h->stream, h->ev_join, 0);
    }
}

void launch_meanflow_prepass(chb_handle_s* h) {
    if (h->g.nx0 != 0) return;
    CHB_LAUNCH(1, 32, 0, h->stream, meanflow_prepass_kernel)(h->V, h->g, h->tab, h->sc);
    h->launches++;
}
#endif
