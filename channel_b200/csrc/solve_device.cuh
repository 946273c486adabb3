// solve_device.cuh - device pieces shared by the y-direction kernels (solve_kernels.cu, rhs_kernel.cu):
// rows of D2vmat / etamat (linsolve_blocking.inc:12-13), wall-BC folding (applybc_0/n,
// dnsdata.f90:458-472) and one row of the banded UL factorisation (LU5decompStep,
// rbparmat_blocking.f90:35-43).
#pragma once
#include "chb_internal.h"

#define SOLVE_K CHB_SOLVE_K   // rows between checkpoints of the UL recurrence

struct Row5 {
    double a[5];
};

// rows of D2vmat / etamat before BC folding                     linsolve_blocking.inc:12-13
__device__ __forceinline__ void build_rows(const DevTables& tab, int iy, double k2, double lam, double ni, Row5& rv,
                                           Row5& re) {
    const int ti = (iy + 1) * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double d0 = __ldg(&tab.d0[ti + j]), d2 = __ldg(&tab.d2[ti + j]), d4 = __ldg(&tab.d4[ti + j]);
        const double OS = ni * (d4 - 2.0 * k2 * d2 + k2 * k2 * d0);  // dnsdata.f90:476
        const double SQ = ni * (d2 - k2 * d0);                       // dnsdata.f90:477
        rv.a[j] = lam * (d2 - k2 * d0) - OS;
        re.a[j] = lam * d0 - SQ;
    }
}

// The same rows from per-substep tables (solve_rows_kernel): with lam fixed during a substep every
// band entry is a polynomial in k2 whose coefficients depend on (iy, j) only,
//   D2vmat: lam*(d2 - k2 d0) - ni*(d4 - 2 k2 d2 + k2^2 d0) = Av + k2*(Bv + k2*Cv)
//   etamat: lam*d0 - ni*(d2 - k2 d0)                       = Ae + k2*Be
// so a row costs 5 x 2 (v) or 5 x 1 (eta) fused multiply-adds instead of ~12 operations per entry.
// rows[(iy+1)*25 + j*5 + {0,1,2,3,4}] = Av, Bv, Cv, Ae, Be
template <int COMP>
__device__ __forceinline__ void build_row_poly(const double* __restrict__ rows, int iy, double k2, Row5& r) {
    const double* t = rows + (size_t)(iy + 1) * 25;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (COMP) r.a[j] = __ldg(&t[j * 5 + 0]) + k2 * (__ldg(&t[j * 5 + 1]) + k2 * __ldg(&t[j * 5 + 2]));
        else r.a[j] = __ldg(&t[j * 5 + 3]) + k2 * __ldg(&t[j * 5 + 4]);
    }
}

// applybc_n / applybc_0 on the rows they touch                  dnsdata.f90:458-472
__device__ __forceinline__ void fold_top1(Row5& r, const double* bcn, const double* bcnp1) {  // row ny-1
    double e = r.a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r.a[j] -= e * bcnp1[j] / bcnp1[4];
    e = r.a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[j] -= e * bcn[j] / bcn[3];
}
__device__ __forceinline__ void fold_top2(Row5& r, const double* bcn) {  // row ny-2
    const double e = r.a[4];
#pragma unroll
    for (int j = 1; j < 4; ++j) r.a[j] -= e * bcn[j - 1] / bcn[3];
}
__device__ __forceinline__ void fold_bot1(Row5& r, const double* bc0, const double* bc0m1) {  // row 1
    double e = r.a[0];
#pragma unroll
    for (int j = 1; j < 5; ++j) r.a[j] -= e * bc0m1[j] / bc0m1[0];
    e = r.a[1];
#pragma unroll
    for (int j = 2; j < 5; ++j) r.a[j] -= e * bc0[j] / bc0[1];
}
__device__ __forceinline__ void fold_bot2(Row5& r, const double* bc0) {  // row 2
    const double e = r.a[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) r.a[j] -= e * bc0[j + 1] / bc0[1];
}

// state of the UL factorisation carried from rows i+1, i+2: their scaled bands -2,-1
struct LUState {
    double l1m2, l1m1;  // row i+1: A(i+1,-2), A(i+1,-1)
    double l2m2, l2m1;  // row i+2
};

// one row of LU5decompStep (rbparmat_blocking.f90:35-43); returns A(i,0)=1/diag, A(i,1), A(i,2)
__device__ __forceinline__ void lu_row(Row5& r, LUState& st, double& inv, double& u1, double& u2) {
    double piv = r.a[4];
    r.a[3] -= piv * st.l2m1;
    r.a[2] -= piv * st.l2m2;
    u2 = piv;
    piv = r.a[3];
    r.a[2] -= piv * st.l1m1;
    r.a[1] -= piv * st.l1m2;
    u1 = piv;
    inv = 1.0 / r.a[2];
    r.a[0] *= inv;
    r.a[1] *= inv;
    st.l2m2 = st.l1m2;
    st.l2m1 = st.l1m1;
    st.l1m2 = r.a[0];
    st.l1m1 = r.a[1];
}

