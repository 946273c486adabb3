// host_tables.cpp - host-side setup that the reference keeps on the CPU and that a Fortran
// driver would pass in through chb_set_tables: grid, compact finite-difference coefficient
// tables and boundary-condition vectors (setup_derivatives dnsdata.f90:241-286,
// setup_boundary_conditions dnsdata.f90:290-308, LUdecomp rbmat.f90:60-76, .bs. rbmat.f90:201-215,
// LU5decompStep rbparmat_blocking.f90:20-51 with npy=1).  Used by the C++/Python host side of
// this repository because no Fortran compiler exists in the build image.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/channel_b200.h"
#include "../../include/channel_b200_host.h"
#include "transpose_index.h"

namespace {

// in-place LU of a dense 5x5 (row-major a[i*5+j]), rbmat.f90:60-76
void LUdecomp(double* A, int HI) {
    for (int i = HI - 1; i >= 1; --i) {
        double piv = 1.0 / A[i * HI + i];
        A[i * HI + i] = piv;
        for (int j = 0; j < i; ++j) A[i * HI + j] *= piv;
        for (int k = 0; k < i; ++k) {
            piv = A[k * HI + i];
            for (int j = 0; j < i; ++j) A[k * HI + j] -= piv * A[i * HI + j];
        }
    }
    A[0] = 1.0 / A[0];
}

// x = A .bs. b, rbmat.f90:201-215
void bs(const double* A, const double* b, double* x, int HI) {
    x[HI - 1] = b[HI - 1] * A[(HI - 1) * HI + HI - 1];
    for (int i = HI - 2; i >= 0; --i) {
        double s = 0.0;
        for (int j = i + 1; j < HI; ++j) s += A[i * HI + j] * x[j];
        x[i] = (b[i] - s) * A[i * HI + i];
    }
    for (int i = 1; i < HI; ++i) {
        double s = 0.0;
        for (int j = 0; j < i; ++j) s += A[i * HI + j] * x[j];
        x[i] = x[i] - s;
    }
}

// rbparmat_blocking.f90:20-51, rows i=0..n-1, bands -2..2 at A[i*5 + j+2]
void LU5decompStep(double* A, int nrows) {
    const int HI1 = nrows - 1;
    A[(HI1 - 2) * 5 + 3] = 0.0; A[(HI1 - 2) * 5 + 4] = 0.0; A[(HI1 - 3) * 5 + 4] = 0.0;
    for (int i = HI1 - 2; i >= 0; --i) {
        for (int k = 2; k >= 1; --k) {
            const double piv = A[i * 5 + k + 2];
            for (int j = -1; j >= -2; --j) A[i * 5 + j + k + 2] -= piv * A[(i + k) * 5 + j + 2];
        }
        const double piv = 1.0 / A[i * 5 + 2];
        A[i * 5 + 2] = piv;
        A[i * 5 + 0] *= piv;
        A[i * 5 + 1] *= piv;
    }
    A[0] = 0.0; A[1] = 0.0; A[5] = 0.0;
}

}  // namespace

extern "C" int chb_host_fft_fit(int n) {  // ffts.f90:78-86
    int j = n;
    if (j <= 0) return 0;
    while (j % 2 == 0) j >>= 1;
    return (j == 1 || j == 3) ? 1 : 0;
}

extern "C" int chb_host_padded_sizes(int nx, int nz, int* nxd, int* nzd) {  // dnsdata.f90:110-113
    int a = 3 * (nx + 1) / 2, b = 3 * nz;
    while (!chb_host_fft_fit(a)) ++a;
    while (!chb_host_fft_fit(b)) ++b;
    *nxd = a;
    *nzd = b;
    return 0;
}

extern "C" int chb_host_setup_tables(int ny, double a, double ymin, double ymax, chb_host_tables* t) {
    if (ny < 8 || !t) return 2;
    const int nyp = ny + 3;
    auto Y = [&](int iy) -> double& { return t->y[iy + 1]; };
    for (int iy = -1; iy <= ny + 1; ++iy)  // dnsdata.f90:153
        Y(iy) = ymin + 0.5 * (ymax - ymin) * (std::tanh(a * (2.0 * (double)iy / (double)ny - 1.0)) / std::tanh(a) + 1.0);
    (void)nyp;
    double M[25], tt[5], h[5];
    for (int iy = 1; iy <= ny - 1; ++iy) {
        double* d0 = t->d0 + (size_t)(iy - 1) * 5;
        double* d1 = t->d1 + (size_t)(iy - 1) * 5;
        double* d2 = t->d2 + (size_t)(iy - 1) * 5;
        double* d4 = t->d4 + (size_t)(iy - 1) * 5;
        for (int j = 0; j < 5; ++j) h[j] = Y(iy - 2 + j) - Y(iy);
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) M[i * 5 + j] = std::pow(h[j], 4.0 - i);   // :247
        LUdecomp(M, 5);
        for (int i = 0; i < 5; ++i) tt[i] = 0.0;
        tt[0] = 24.0;
        bs(M, tt, d4, 5);                                                                                  // :249
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j)
            M[i * 5 + j] = (5.0 - i) * (6.0 - i) * (7.0 - i) * (8.0 - i) * std::pow(h[j], 4.0 - i);        // :250
        LUdecomp(M, 5);
        for (int i = 0; i < 5; ++i) {                                                                      // :251
            double s = 0.0;
            for (int j = 0; j < 5; ++j) s += d4[j] * std::pow(h[j], 8.0 - i);
            tt[i] = s;
        }
        bs(M, tt, d0, 5);                                                                                  // :252
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) M[i * 5 + j] = std::pow(h[j], 4.0 - i);   // :253
        LUdecomp(M, 5);
        for (int i = 0; i < 5; ++i) tt[i] = 0.0;
        for (int i = 0; i < 3; ++i) {                                                                      // :254
            double s = 0.0;
            for (int j = 0; j < 5; ++j) s += d0[j] * (4.0 - i) * (3.0 - i) * std::pow(h[j], 2.0 - i);
            tt[i] = s;
        }
        bs(M, tt, d2, 5);
        for (int i = 0; i < 5; ++i) tt[i] = 0.0;
        for (int i = 0; i < 4; ++i) {                                                                      // :256
            double s = 0.0;
            for (int j = 0; j < 5; ++j) s += d0[j] * (4.0 - i) * std::pow(h[j], 3.0 - i);
            tt[i] = s;
        }
        bs(M, tt, d1, 5);
    }
    auto wall = [&](int node0, int base, double* a1, double* a2) {
        for (int j = 0; j < 5; ++j) h[j] = Y(node0 + j) - Y(base);
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) M[i * 5 + j] = std::pow(h[j], 4.0 - i);
        LUdecomp(M, 5);
        for (int i = 0; i < 5; ++i) tt[i] = 0.0;
        tt[3] = 1.0;
        bs(M, tt, a1, 5);
        for (int i = 0; i < 5; ++i) tt[i] = 0.0;
        tt[2] = 2.0;
        bs(M, tt, a2, 5);
    };
    wall(-1, 0, t->d140, t->d240);            // :260-262
    wall(-1, -1, t->d14m1, t->d24m1);         // :263-265
    wall(ny - 3, ny, t->d14n, t->d24n);       // :269-271
    wall(ny - 3, ny + 1, t->d14np1, t->d24np1);  // :272-274
    double d040[5] = {0, 1, 0, 0, 0}, d04n[5] = {0, 0, 0, 1, 0};  // :266,275
    // D0mat                                   :277,284
    for (int i = 0; i < (ny + 1) * 5; ++i) t->D0mat[i] = 0.0;
    memcpy(t->D0mat, t->d0, sizeof(double) * 5 * (ny - 1));
    LU5decompStep(t->D0mat, ny + 1);
    // setup_boundary_conditions               :290-308 (full channel)
    memcpy(t->v0bc, d040, sizeof(d040)); memcpy(t->v0m1bc, t->d140, sizeof(d040)); memcpy(t->eta0bc, d040, sizeof(d040));
    memcpy(t->eta0m1bc, t->d4, sizeof(d040));  // der(1)%d4
    {
        const double e = t->v0bc[0];
        for (int j = 1; j < 5; ++j) t->v0bc[j] -= e * t->v0m1bc[j] / t->v0m1bc[0];
        const double f = t->eta0bc[0];
        for (int j = 1; j < 5; ++j) t->eta0bc[j] -= f * t->eta0m1bc[j] / t->eta0m1bc[0];
    }
    memcpy(t->vnbc, d04n, sizeof(d04n)); memcpy(t->vnp1bc, t->d14n, sizeof(d04n)); memcpy(t->etanbc, d04n, sizeof(d04n));
    memcpy(t->etanp1bc, t->d4 + (size_t)(ny - 2) * 5, sizeof(d04n));  // der(ny-1)%d4
    {
        const double e = t->vnbc[4];
        for (int j = 0; j < 4; ++j) t->vnbc[j] -= e * t->vnp1bc[j] / t->vnp1bc[4];
        const double f = t->etanbc[4];
        for (int j = 0; j < 4; ++j) t->etanbc[j] -= f * t->etanp1bc[j] / t->etanp1bc[4];
    }
    return 0;
}

extern "C" int chb_host_apply_tables(chb_handle h, const chb_host_tables* t) {
    return chb_set_tables(h, t->y, t->d0, t->d1, t->d2, t->d4, t->d140, t->d14m1, t->d240, t->d24m1, t->d14n, t->d14np1,
                          t->d24n, t->d24np1, t->v0bc, t->v0m1bc, t->vnbc, t->vnp1bc, t->eta0bc, t->eta0m1bc, t->etanbc,
                          t->etanp1bc, t->D0mat);
}

extern "C" int chb_host_decomposition(int nx, int nzd, int nranks, int rank, int* nx0, int* nxN, int* nz0, int* nzN) {
    if (nranks < 1 || rank < 0 || rank >= nranks || (nx + 1) % nranks != 0 || nzd % nranks != 0) return 2;
    chb_decompose(nx + 1, nzd, nranks, rank, nx0, nxN, nz0, nzN);
    return 0;
}

extern "C" long long chb_host_transpose_index(int peer, int ncomp, int comp, int nplanes, int plane, int nzB, int izl,
                                              int nxB, int ixl) {
    return (long long)chb_buf_index(peer, ncomp, comp, nplanes, plane, nzB, izl, nxB, ixl);
}
