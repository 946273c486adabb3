/* channel_oracle_c.c - CPU restatement (C99 + OpenMP) of the per-timestep hot path of
 * davecats/channel, plane by plane exactly as the reference does it.
 *
 * TEST INFRASTRUCTURE ONLY.  It is the checker of tests/ and __graft_entry__.smoke() and the
 * timed CPU baseline of bench.py (cpu_baseline / --impl reference).  Nothing under
 * channel_b200/ (the product) links, loads or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path and
 * cannot be built in this image (no Fortran compiler, no MPI, no FFTW), see DESIGN.md.  This
 * file is pinned by (1) analytic known answers, (2) agreement with the independent numpy
 * restatement oracle/channel_oracle.py (pocketfft FFTs, all planes at once), both checked in
 * tests/test_oracle_*.py.
 *
 * What follows what (all file:line into the reference tree):
 *   co_create            read_dnsin sizes dnsdata.f90:110-124, fftFIT ffts.f90:78-86,
 *                        init_memory dnsdata.f90:129-158 (grid, izd, ialfa, ibeta, k2)
 *   setup_derivatives    dnsdata.f90:241-286 with LUdecomp rbmat.f90:60-76, .bs. rbmat.f90:201-215
 *   setup_bc             dnsdata.f90:290-308; applybc_0/n dnsdata.f90:458-472
 *   LU5decompStep, LeftLU5divStep1/2   rbparmat_blocking.f90:20-100 (npy=1: first=last=.TRUE.)
 *   convolutions         dnsdata.f90:487-602 (non-ibm, non-convvel, blocking branches);
 *                        zTOx/xTOz mpi_transpose.f90:50-117 with one rank = plain transposes
 *   buildrhs             dnsdata.f90:611-673 (5-slot VVdz ring imod, 3-slot memrhs ring, delayed
 *                        write-back into V), body-force ghost rows :616-629
 *   linsolve             linsolve_blocking.inc:3-107 (incl. mean mode, CPI, inline vetaTOuvw)
 *   COMPLEXderiv         dnsdata.f90:339-373;  yintegr dnsdata.f90:312-324
 *   cfl_prepass/outstats channel.f90:95-116, dnsdata.f90:853-880
 *   coriolis force       body_forces/coriolis/coriolis.inc:29-41
 *
 * FFTW 3.x (un-vendored third-party dependency of the reference, ffts.f90:56-75) is restated
 * by a Stockham autosort mixed-radix (4,2,3) complex FFT; the real transforms of logical
 * length 2*nxd are done as complex transforms of length nxd plus the standard split/merge
 * pass.  Conventions as in ffts.f90: IFT sign +, FFT sign -, c2r sign + (imaginary parts of
 * the DC and Nyquist inputs ignored), r2c sign -, all unnormalised.
 *
 * Memory layout is the reference's: V(iy,iz,ix,c) with iy fastest (dnsdata.f90:132),
 * VVdz(nzd,nxB,6,6), VVdx(nxd+1,nzB,6,6) aliased with rVVdx(2nxd+2,nzB,6,6) (ffts.f90:56-64).
 * Deviations that do not change any result: OpenMP threads over lines/columns stand in for MPI
 * ranks over pencils; linsolve's two iz-loops are merged per column (one matrix pair per thread).
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#else
static double omp_get_wtime(void) { return 0.0; }
static int omp_get_max_threads(void) { return 1; }
static int omp_get_thread_num(void) { return 0; }
#endif

typedef double _Complex cplx;
#define PI_REF 3.1415926535897932384626433832795028841971 /* dnsdata.f90:26 */

/* ------------------------------------------------------------------------------------------
 * FFT: Stockham autosort, radices 4,2,3; tw[k] = exp(+2 pi i k / n)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n, nf, fac[40];
    cplx* tw;
} fftplan;

static int fft_fit(int n) { /* ffts.f90:78-86 */
    int j = n;
    while (j % 2 == 0) j /= 2;
    return j == 1 || j == 3;
}

static void plan_init(fftplan* p, int n) {
    p->n = n;
    p->nf = 0;
    int r = n;
    while (r % 4 == 0) { p->fac[p->nf++] = 4; r /= 4; }
    while (r % 2 == 0) { p->fac[p->nf++] = 2; r /= 2; }
    while (r % 3 == 0) { p->fac[p->nf++] = 3; r /= 3; }
    if (r != 1) { fprintf(stderr, "oracle fft: unsupported size %d\n", n); abort(); }
    p->tw = (cplx*)malloc(sizeof(cplx) * (size_t)n);
    for (int k = 0; k < n; ++k) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        p->tw[k] = (double)cosl(a) + I * (double)sinl(a);
    }
}

/* in-place transform of x[0..n) using work[0..n); sign = +1 or -1 */
static void fft_line(const fftplan* p, cplx* x, cplx* work, int sign) {
    const int N = p->n;
    cplx* a = x;
    cplx* b = work;
    int n = N, s = 1;
    for (int f = 0; f < p->nf; ++f) {
        const int r = p->fac[f];
        const int m = n / r;
        if (r == 2) {
            for (int pp = 0; pp < m; ++pp) {
                cplx w = p->tw[(size_t)pp * s];
                if (sign < 0) w = conj(w);
                for (int q = 0; q < s; ++q) {
                    const cplx u = a[q + s * pp], v = a[q + s * (pp + m)];
                    b[q + s * (2 * pp)] = u + v;
                    b[q + s * (2 * pp + 1)] = (u - v) * w;
                }
            }
        } else if (r == 4) {
            for (int pp = 0; pp < m; ++pp) {
                cplx w1 = p->tw[(size_t)pp * s], w2 = p->tw[(size_t)2 * pp * s], w3 = p->tw[(size_t)3 * pp * s];
                if (sign < 0) { w1 = conj(w1); w2 = conj(w2); w3 = conj(w3); }
                for (int q = 0; q < s; ++q) {
                    const cplx x0 = a[q + s * pp], x1 = a[q + s * (pp + m)];
                    const cplx x2 = a[q + s * (pp + 2 * m)], x3 = a[q + s * (pp + 3 * m)];
                    const cplx t0 = x0 + x2, t1 = x0 - x2, t2 = x1 + x3;
                    const cplx d = x1 - x3;
                    const cplx t3 = (sign > 0) ? (-cimag(d) + I * creal(d)) : (cimag(d) - I * creal(d));
                    b[q + s * (4 * pp)] = t0 + t2;
                    b[q + s * (4 * pp + 1)] = (t1 + t3) * w1;
                    b[q + s * (4 * pp + 2)] = (t0 - t2) * w2;
                    b[q + s * (4 * pp + 3)] = (t1 - t3) * w3;
                }
            }
        } else { /* r == 3 */
            const double s3 = 0.86602540378443864676372317075294;
            for (int pp = 0; pp < m; ++pp) {
                cplx w1 = p->tw[(size_t)pp * s], w2 = p->tw[(size_t)2 * pp * s];
                if (sign < 0) { w1 = conj(w1); w2 = conj(w2); }
                for (int q = 0; q < s; ++q) {
                    const cplx x0 = a[q + s * pp], x1 = a[q + s * (pp + m)], x2 = a[q + s * (pp + 2 * m)];
                    const cplx t = x1 + x2, d = (x1 - x2) * s3;
                    const cplx mm = x0 - 0.5 * t;
                    const cplx rr = (sign > 0) ? (-cimag(d) + I * creal(d)) : (cimag(d) - I * creal(d));
                    b[q + s * (3 * pp)] = x0 + t;
                    b[q + s * (3 * pp + 1)] = (mm + rr) * w1;
                    b[q + s * (3 * pp + 2)] = (mm - rr) * w2;
                }
            }
        }
        cplx* t = a; a = b; b = t;
        n = m;
        s *= r;
    }
    if (a != x) memcpy(x, a, sizeof(cplx) * (size_t)N);
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int nx, ny, nz, nxd, nzd, nzt, nyp;
    double alfa0, beta0, ni, a, ymin, ymax;
    double dx, dz, factor;
    double meanpx, meanpz, meanflowx, meanflowz, gamma, u0, uN, deltat, cflmax, time;
    int CPI, CPI_type;
    double cfl, fr[3], corrpx, corrpz;
    double *y, *dy;           /* index iy+1 */
    double *d0, *d1, *d2, *d4; /* [(iy+1)*5 + j+2], rows iy=1..ny-1 */
    double d140[5], d14m1[5], d240[5], d24m1[5], d14n[5], d14np1[5], d24n[5], d24np1[5], d040[5], d04n[5];
    double v0bc[5], v0m1bc[5], vnbc[5], vnp1bc[5], eta0bc[5], eta0m1bc[5], etanbc[5], etanp1bc[5];
    double* D0mat;            /* [(ny+1)*5] row i <-> iy=i+1 */
    int* izd;                 /* index iz+nz */
    cplx* V;                  /* V(iy,iz,ix,c) */
    cplx* F;                  /* body force, same layout, or NULL */
    cplx* oldrhs;             /* oldrhs(1:ny-1,-nz:nz,0:nx) of {eta,d2v} */
    cplx* memrhs;             /* memrhs(0:2,-nz:nz,0:nx) of {eta,d2v} */
    cplx* VVdz;               /* VVdz(nzd,nx+1,6,5 used slots) */
    cplx* VVdx;               /* VVdx(nxd+1,nzd,6) (one slot is enough: it is transient) */
    fftplan pz, px;
    cplx* wh;                 /* exp(+i pi k / nxd), k=0..nxd */
    int nthreads;
    cplx** work;              /* per-thread FFT work, 3*max(nzd,nxd+1) */
    /* body force: 1 = coriolis (body_forces/coriolis/coriolis.inc), 2 = am_f1, 3 = am_butterfly */
    int bodyforce;
    double am_amp;
    int am_iz_f;
    double omega2;
    int iz_thr;
    double y_thr_bot, y_thr_top;
    /* sample timers */
    double t_conv, t_rhs, t_solve;
} co_state;

#define IDXV(st, c, ix, izp, iyp) ((((size_t)(c) * ((st)->nx + 1) + (ix)) * (st)->nzt + (izp)) * (st)->nyp + (iyp))
#define YY(st, iy) ((st)->y[(iy) + 1])
#define DER(tab, iy, j) ((tab)[((iy) + 1) * 5 + (j) + 2])

/* rbmat.f90:60-76 */
static void LUdecomp(double A[5][5]) {
    for (int i = 4; i >= 1; --i) {
        double piv = 1.0 / A[i][i];
        A[i][i] = piv;
        for (int j = 0; j < i; ++j) A[i][j] = A[i][j] * piv;
        for (int k = 0; k < i; ++k) {
            piv = A[k][i];
            for (int j = 0; j < i; ++j) A[k][j] = A[k][j] - piv * A[i][j];
        }
    }
    A[0][0] = 1.0 / A[0][0];
}
/* rbmat.f90:201-215 */
static void bs(double A[5][5], const double b[5], double x[5]) {
    x[4] = b[4] * A[4][4];
    for (int i = 3; i >= 0; --i) {
        double s = 0.0;
        for (int j = i + 1; j < 5; ++j) s += A[i][j] * x[j];
        x[i] = (b[i] - s) * A[i][i];
    }
    for (int i = 1; i < 5; ++i) {
        double s = 0.0;
        for (int j = 0; j < i; ++j) s += A[i][j] * x[j];
        x[i] = x[i] - s;
    }
}

/* rbparmat_blocking.f90:20-51; A rows 0..nrows-1, bands at A[i*5+j+2] */
static void LU5decompStep(double* A, int nrows) {
    const int HI1 = nrows - 1;
    A[(HI1 - 2) * 5 + 3] = 0; A[(HI1 - 2) * 5 + 4] = 0; A[(HI1 - 3) * 5 + 4] = 0;
    for (int i = HI1 - 2; i >= 0; --i) {
        for (int k = 2; k >= 1; --k) {
            const double piv = A[i * 5 + k + 2];
            for (int j = -1; j >= -2; --j) A[i * 5 + j + k + 2] = A[i * 5 + j + k + 2] - piv * A[(i + k) * 5 + j + 2];
        }
        const double piv = 1.0 / A[i * 5 + 2];
        A[i * 5 + 2] = piv;
        A[i * 5 + 0] *= piv;
        A[i * 5 + 1] *= piv;
    }
    A[0] = 0; A[1] = 0; A[5] = 0;
}
/* x(-2:HI1) <-> x[i+2]; rbparmat_blocking.f90:57-77 */
static void LeftLU5divStep1(cplx* x, const double* A, int nrows) {
    const int HI1 = nrows - 1;
    for (int i = HI1 - 2; i >= 0; --i)
        x[i + 2] = (x[i + 2] - (A[i * 5 + 3] * x[i + 3] + A[i * 5 + 4] * x[i + 4])) * A[i * 5 + 2];
}
/* rbparmat_blocking.f90:82-100 */
static void LeftLU5divStep2(const double* A, cplx* b, int nrows) {
    const int HI1 = nrows - 1;
    for (int i = 0; i <= HI1; ++i) b[i + 2] = b[i + 2] - (A[i * 5 + 0] * b[i] + A[i * 5 + 1] * b[i + 1]);
}

static void wall_stencils(const co_state* st, int node0, int base, double* a1, double* a2) {
    double M[5][5], t[5];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) M[i][j] = pow(YY(st, node0 + j) - YY(st, base), 4.0 - i);
    LUdecomp(M);
    memset(t, 0, sizeof t); t[3] = 1.0; bs(M, t, a1);
    memset(t, 0, sizeof t); t[2] = 2.0; bs(M, t, a2);
}

static void setup_derivatives(co_state* st) { /* dnsdata.f90:241-286 */
    const int ny = st->ny;
    for (int iy = 1; iy <= ny - 1; ++iy) {
        double M[5][5], t[5], h[5], d4[5], d0[5], d2[5], d1[5];
        for (int j = 0; j < 5; ++j) h[j] = YY(st, iy - 2 + j) - YY(st, iy);
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) M[i][j] = pow(h[j], 4.0 - i);
        LUdecomp(M);
        memset(t, 0, sizeof t); t[0] = 24; bs(M, t, d4);
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) M[i][j] = (5.0 - i) * (6.0 - i) * (7.0 - i) * (8.0 - i) * pow(h[j], 4.0 - i);
        LUdecomp(M);
        for (int i = 0; i < 5; ++i) { double s = 0; for (int j = 0; j < 5; ++j) s += d4[j] * pow(h[j], 8.0 - i); t[i] = s; }
        bs(M, t, d0);
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) M[i][j] = pow(h[j], 4.0 - i);
        LUdecomp(M);
        memset(t, 0, sizeof t);
        for (int i = 0; i <= 2; ++i) { double s = 0; for (int j = 0; j < 5; ++j) s += d0[j] * (4.0 - i) * (3.0 - i) * pow(h[j], 2.0 - i); t[i] = s; }
        bs(M, t, d2);
        memset(t, 0, sizeof t);
        for (int i = 0; i <= 3; ++i) { double s = 0; for (int j = 0; j < 5; ++j) s += d0[j] * (4.0 - i) * pow(h[j], 3.0 - i); t[i] = s; }
        bs(M, t, d1);
        for (int j = 0; j < 5; ++j) {
            st->d0[(iy + 1) * 5 + j] = d0[j]; st->d1[(iy + 1) * 5 + j] = d1[j];
            st->d2[(iy + 1) * 5 + j] = d2[j]; st->d4[(iy + 1) * 5 + j] = d4[j];
        }
    }
    wall_stencils(st, -1, 0, st->d140, st->d240);
    wall_stencils(st, -1, -1, st->d14m1, st->d24m1);
    memset(st->d040, 0, sizeof st->d040); st->d040[1] = 1;
    wall_stencils(st, ny - 3, ny, st->d14n, st->d24n);
    wall_stencils(st, ny - 3, ny + 1, st->d14np1, st->d24np1);
    memset(st->d04n, 0, sizeof st->d04n); st->d04n[3] = 1;
    memset(st->D0mat, 0, sizeof(double) * 5 * (ny + 1)); /* rows ny, ny+1 zero */
    for (int iy = 1; iy <= ny - 1; ++iy) memcpy(st->D0mat + (iy - 1) * 5, st->d0 + (iy + 1) * 5, sizeof(double) * 5);
    LU5decompStep(st->D0mat, ny + 1);
}

static void setup_bc(co_state* st) { /* dnsdata.f90:290-308 */
    const int ny = st->ny;
    memcpy(st->v0bc, st->d040, 40); memcpy(st->v0m1bc, st->d140, 40); memcpy(st->eta0bc, st->d040, 40);
    memcpy(st->eta0m1bc, st->d4 + (1 + 1) * 5, 40);
    { const double e = st->v0bc[0]; for (int j = 1; j < 5; ++j) st->v0bc[j] = st->v0bc[j] - e * st->v0m1bc[j] / st->v0m1bc[0]; }
    { const double e = st->eta0bc[0]; for (int j = 1; j < 5; ++j) st->eta0bc[j] = st->eta0bc[j] - e * st->eta0m1bc[j] / st->eta0m1bc[0]; }
    memcpy(st->vnbc, st->d04n, 40); memcpy(st->vnp1bc, st->d14n, 40); memcpy(st->etanbc, st->d04n, 40);
    memcpy(st->etanp1bc, st->d4 + (ny - 1 + 1) * 5, 40);
    { const double e = st->vnbc[4]; for (int j = 0; j < 4; ++j) st->vnbc[j] = st->vnbc[j] - e * st->vnp1bc[j] / st->vnp1bc[4]; }
    { const double e = st->etanbc[4]; for (int j = 0; j < 4; ++j) st->etanbc[j] = st->etanbc[j] - e * st->etanp1bc[j] / st->etanp1bc[4]; }
}

/* dnsdata.f90:458-472 on a matrix with rows iy=1.. at A[(iy-1)*5 + j+2] */
static void applybc_0(double* A, const double* bc0, const double* bc0m1) {
    double e = A[0];
    for (int j = 1; j < 5; ++j) A[j] = A[j] - e * bc0m1[j] / bc0m1[0];
    e = A[1];
    for (int j = 2; j < 5; ++j) A[j] = A[j] - e * bc0[j] / bc0[1];
    e = A[5 + 0];
    for (int j = 1; j < 4; ++j) A[5 + j] = A[5 + j] - e * bc0[j + 1] / bc0[1];
}
static void applybc_n(double* A, int ny, const double* bcn, const double* bcnp1) {
    double* r1 = A + (ny - 2) * 5; /* iy = ny-1 */
    double* r2 = A + (ny - 3) * 5; /* iy = ny-2 */
    double e = r1[4];
    for (int j = 0; j < 4; ++j) r1[j] = r1[j] - e * bcnp1[j] / bcnp1[4];
    e = r1[3];
    for (int j = 0; j < 3; ++j) r1[j] = r1[j] - e * bcn[j] / bcn[3];
    e = r2[4];
    for (int j = 1; j < 4; ++j) r2[j] = r2[j] - e * bcn[j - 1] / bcn[3];
}

/* dnsdata.f90:312-324; f indexed iy+1 with stride st_ (in doubles) */
static double yintegr(const co_state* st, const double* f, size_t stride) {
    double II = 0.0;
    for (int iy = 1; iy <= st->ny - 1; iy += 2) {
        const double yp1 = YY(st, iy + 1) - YY(st, iy), ym1 = YY(st, iy - 1) - YY(st, iy);
        const double a1 = -1.0 / 3.0 * ym1 + 1.0 / 6.0 * yp1 + 1.0 / 6.0 * yp1 * yp1 / ym1;
        const double a3 = +1.0 / 3.0 * yp1 - 1.0 / 6.0 * ym1 - 1.0 / 6.0 * ym1 * ym1 / yp1;
        const double a2 = yp1 - ym1 - a1 - a3;
        II = II + a1 * f[(size_t)(iy) * stride] + a2 * f[(size_t)(iy + 1) * stride] + a3 * f[(size_t)(iy + 2) * stride];
    }
    return II;
}

/* ------------------------------------------------------------------------------------------ */
int co_padded_sizes(int nx, int nz, int* nxd, int* nzd) { /* dnsdata.f90:110-113 */
    int a = 3 * (nx + 1) / 2, b = 3 * nz;
    while (!fft_fit(a)) ++a;
    while (!fft_fit(b)) ++b;
    *nxd = a; *nzd = b;
    return 0;
}

void co_destroy(co_state* st);

co_state* co_create(int nx, int ny, int nz, double alfa0, double beta0, double ni, double a, double ymin, double ymax) {
    co_state* st = (co_state*)calloc(1, sizeof(co_state));
    st->nx = nx; st->ny = ny; st->nz = nz;
    co_padded_sizes(nx, nz, &st->nxd, &st->nzd);
    st->nzt = 2 * nz + 1; st->nyp = ny + 3;
    st->alfa0 = alfa0; st->beta0 = beta0; st->ni = ni; st->a = a; st->ymin = ymin; st->ymax = ymax;
    st->dx = PI_REF / (alfa0 * st->nxd); st->dz = 2.0 * PI_REF / (beta0 * st->nzd);
    st->factor = 1.0 / (2.0 * st->nxd * st->nzd); /* dnsdata.f90:124 */
    st->y = (double*)calloc(st->nyp, 8); st->dy = (double*)calloc(st->nyp, 8);
    for (int iy = -1; iy <= ny + 1; ++iy)
        st->y[iy + 1] = ymin + 0.5 * (ymax - ymin) * (tanh(a * (2.0 * (double)iy / (double)ny - 1)) / tanh(a) + 1);
    for (int iy = 1; iy <= ny - 1; ++iy) st->dy[iy + 1] = 0.5 * (YY(st, iy + 1) - YY(st, iy - 1));
    st->d0 = (double*)calloc((size_t)st->nyp * 5, 8); st->d1 = (double*)calloc((size_t)st->nyp * 5, 8);
    st->d2 = (double*)calloc((size_t)st->nyp * 5, 8); st->d4 = (double*)calloc((size_t)st->nyp * 5, 8);
    st->D0mat = (double*)calloc((size_t)(ny + 1) * 5, 8);
    st->izd = (int*)malloc(sizeof(int) * st->nzt);
    for (int iz = -nz; iz <= nz; ++iz) st->izd[iz + nz] = iz >= 0 ? iz : st->nzd + iz;
    setup_derivatives(st);
    setup_bc(st);
    const size_t nV = (size_t)3 * (nx + 1) * st->nzt * st->nyp;
    st->V = (cplx*)calloc(nV, sizeof(cplx));
    st->oldrhs = (cplx*)calloc((size_t)2 * (nx + 1) * st->nzt * (ny - 1), sizeof(cplx));
    st->memrhs = (cplx*)calloc((size_t)2 * 3 * (nx + 1) * st->nzt, sizeof(cplx));
    st->VVdz = (cplx*)calloc((size_t)5 * 6 * (nx + 1) * st->nzd, sizeof(cplx));
    st->VVdx = (cplx*)calloc((size_t)6 * st->nzd * (st->nxd + 1), sizeof(cplx));
    if (!st->V || !st->oldrhs || !st->memrhs || !st->VVdz || !st->VVdx) { co_destroy(st); return NULL; }
    plan_init(&st->pz, st->nzd);
    plan_init(&st->px, st->nxd);
    st->wh = (cplx*)malloc(sizeof(cplx) * (st->nxd + 1));
    for (int k = 0; k <= st->nxd; ++k) {
        long double ang = 3.14159265358979323846264338327950288L * (long double)k / (long double)st->nxd;
        st->wh[k] = (double)cosl(ang) + I * (double)sinl(ang);
    }
    st->nthreads = omp_get_max_threads();
    st->work = (cplx**)malloc(sizeof(cplx*) * st->nthreads);
    const size_t wl = (size_t)3 * ((st->nzd > st->nxd + 1) ? st->nzd : st->nxd + 1);
    for (int t = 0; t < st->nthreads; ++t) st->work[t] = (cplx*)malloc(sizeof(cplx) * wl);
    st->CPI = 0; st->CPI_type = 0; st->gamma = 0; st->deltat = 0; st->cflmax = 0;
    return st;
}

void co_destroy(co_state* st) {
    if (!st) return;
    free(st->y); free(st->dy); free(st->d0); free(st->d1); free(st->d2); free(st->d4); free(st->D0mat); free(st->izd);
    free(st->V); free(st->F); free(st->oldrhs); free(st->memrhs); free(st->VVdz); free(st->VVdx);
    free(st->pz.tw); free(st->px.tw); free(st->wh);
    if (st->work) { for (int t = 0; t < st->nthreads; ++t) free(st->work[t]); free(st->work); }
    free(st);
}

void co_set_params(co_state* st, double meanpx, double meanpz, double meanflowx, double meanflowz, int CPI,
                   int CPI_type, double gamma, double u0, double uN, double deltat, double cflmax, double time) {
    st->meanpx = meanpx; st->meanpz = meanpz; st->meanflowx = meanflowx; st->meanflowz = meanflowz;
    st->CPI = CPI; st->CPI_type = CPI_type; st->gamma = gamma; st->u0 = u0; st->uN = uN;
    st->deltat = deltat; st->cflmax = cflmax; st->time = time;
}
/* Number of OpenMP threads for handles created afterwards (and their per-thread work arrays).  Launchers such as
 * torchrun export OMP_NUM_THREADS=1; the CPU baseline of bench.py must not inherit that. */
void co_global_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
void co_set_threads(co_state* st, int n) {
#ifdef _OPENMP
    if (n >= 1 && n <= st->nthreads) omp_set_num_threads(n);
#else
    (void)st; (void)n;
#endif
}
int co_threads(co_state* st) { (void)st; return omp_get_max_threads(); }
cplx* co_V(co_state* st) { return st->V; }
cplx* co_F(co_state* st) { return st->F; }
cplx* co_oldrhs(co_state* st) { return st->oldrhs; }
void co_sizes(co_state* st, int* nxd, int* nzd) { *nxd = st->nxd; *nzd = st->nzd; }
/* tables, for comparison with the numpy oracle: which = 0..3 -> d0,d1,d2,d4 [(ny+3)*5]; 4 -> D0mat;
 * 5 -> y; 10.. -> wall stencils / BC vectors */
const double* co_table(co_state* st, int which) {
    switch (which) {
        case 0: return st->d0; case 1: return st->d1; case 2: return st->d2; case 3: return st->d4;
        case 4: return st->D0mat; case 5: return st->y;
        case 10: return st->d140; case 11: return st->d14m1; case 12: return st->d240; case 13: return st->d24m1;
        case 14: return st->d14n; case 15: return st->d14np1; case 16: return st->d24n; case 17: return st->d24np1;
        case 20: return st->v0bc; case 21: return st->v0m1bc; case 22: return st->vnbc; case 23: return st->vnp1bc;
        case 24: return st->eta0bc; case 25: return st->eta0m1bc; case 26: return st->etanbc; case 27: return st->etanp1bc;
    }
    return NULL;
}
void co_get_scalars(co_state* st, double* out) { /* cfl, fr[3], corrpx, corrpz, meanpx, meanpz, deltat, time */
    out[0] = st->cfl; out[1] = st->fr[0]; out[2] = st->fr[1]; out[3] = st->fr[2]; out[4] = st->corrpx;
    out[5] = st->corrpz; out[6] = st->meanpx; out[7] = st->meanpz; out[8] = st->deltat; out[9] = st->time;
}

/* batched complex FFT of lines, for the FFT unit tests: data[nlines][n] in place */
void co_fft_lines(int n, int nlines, int sign, cplx* data) {
    fftplan p;
    plan_init(&p, n);
#pragma omp parallel
    {
        cplx* w = (cplx*)malloc(sizeof(cplx) * n);
#pragma omp for
        for (int l = 0; l < nlines; ++l) fft_line(&p, data + (size_t)l * n, w, sign);
        free(w);
    }
    free(p.tw);
}

/* c2r (RFT, ffts.f90:72-73,99-102) of one line: in nxd+1 complex, out 2*nxd reals in place */
static void rft_line(const co_state* st, cplx* line, cplx* work) {
    const int M = st->nxd;
    cplx* Z = work;           /* M */
    cplx* W2 = work + M;      /* M */
    const double x0 = creal(line[0]), xM = creal(line[M]);
    Z[0] = (x0 + xM) + I * (x0 - xM);
    for (int k = 1; k < M; ++k) {
        const cplx a = line[k], b = conj(line[M - k]);
        Z[k] = (a + b) + I * st->wh[k] * (a - b);
    }
    fft_line(&st->px, Z, W2, +1);
    double* r = (double*)line;
    for (int m = 0; m < M; ++m) { r[2 * m] = creal(Z[m]); r[2 * m + 1] = cimag(Z[m]); }
}
/* r2c (HFT, ffts.f90:74-75,104-108): in 2*nxd reals, out nxd+1 complex in place */
static void hft_line(const co_state* st, cplx* line, cplx* work) {
    const int M = st->nxd;
    cplx* Z = work;
    cplx* W2 = work + M;
    const double* r = (const double*)line;
    for (int m = 0; m < M; ++m) Z[m] = r[2 * m] + I * r[2 * m + 1];
    fft_line(&st->px, Z, W2, -1);
    for (int k = 0; k <= M; ++k) {
        const cplx z = Z[k == M ? 0 : k], zm = conj(Z[k == 0 ? 0 : M - k]);
        line[k] = 0.5 * (z + zm) - 0.5 * I * conj(st->wh[k]) * (z - zm);
    }
}

#define VVDZ(st, slot, k, ix) ((st)->VVdz + ((((size_t)(slot) * 6 + (k)) * ((st)->nx + 1)) + (ix)) * (st)->nzd)
#define VVDX(st, k, izd_) ((st)->VVdx + (((size_t)(k) * (st)->nzd) + (izd_)) * ((st)->nxd + 1))

/* dnsdata.f90:487-602 for plane iy into ring slot `slot` */
static void convolutions(co_state* st, int iy, int slot, int compute_cfl, int products) {
    const int nx = st->nx, nz = st->nz, nxd = st->nxd, nzd = st->nzd;
    const int iyp = iy + 1;
    double cflmax = 0.0;
#pragma omp parallel
    {
        cplx* work = st->work[omp_get_thread_num()];
        /* :504-510  pad + IFT */
#pragma omp for collapse(2)
        for (int c = 0; c < 3; ++c)
            for (int ix = 0; ix <= nx; ++ix) {
                cplx* line = VVDZ(st, slot, c, ix);
                for (int iz = 0; iz <= nz; ++iz) line[iz] = st->V[IDXV(st, c, ix, iz + nz, iyp)];
                for (int k = nz + 1; k < nzd - nz; ++k) line[k] = 0;
                for (int iz = -nz; iz <= -1; ++iz) line[nzd + iz] = st->V[IDXV(st, c, ix, iz + nz, iyp)];
                fft_line(&st->pz, line, work, +1);
            }
        /* :533 zTOx (one rank: transpose), :535 zero-pad + RFT */
#pragma omp for collapse(2)
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < nzd; ++k) {
                cplx* xl = VVDX(st, c, k);
                for (int ix = 0; ix <= nx; ++ix) xl[ix] = VVDZ(st, slot, c, ix)[k];
                for (int ix = nx + 1; ix <= nxd; ++ix) xl[ix] = 0;
                rft_line(st, xl, work);
            }
        /* :552-556 cfl, :581-584 products, :586 HFT */
        if (compute_cfl && iy >= 1 && iy <= st->ny - 1) {
            const double dyi = st->dy[iyp];
#pragma omp for reduction(max : cflmax)
            for (int k = 0; k < nzd; ++k) {
                const double* ru = (const double*)VVDX(st, 0, k);
                const double* rv = (const double*)VVDX(st, 1, k);
                const double* rw = (const double*)VVDX(st, 2, k);
                for (int n = 0; n < 2 * nxd; ++n) {
                    const double v = fabs(ru[n]) / st->dx + fabs(rv[n]) / dyi + fabs(rw[n]) / st->dz;
                    if (v > cflmax) cflmax = v;
                }
            }
        }
        if (products) {
            const double f = st->factor;
#pragma omp for
            for (int k = 0; k < nzd; ++k) {
                double* r1 = (double*)VVDX(st, 0, k); double* r2 = (double*)VVDX(st, 1, k); double* r3 = (double*)VVDX(st, 2, k);
                double* r4 = (double*)VVDX(st, 3, k); double* r5 = (double*)VVDX(st, 4, k); double* r6 = (double*)VVDX(st, 5, k);
                for (int n = 0; n < 2 * nxd; ++n) {
                    const double u = r1[n], v = r2[n], w = r3[n];
                    r4[n] = u * v * f; r5[n] = v * w * f; r6[n] = u * w * f;
                    r1[n] = u * u * f; r2[n] = v * v * f; r3[n] = w * w * f;
                }
                for (int c = 0; c < 6; ++c) hft_line(st, VVDX(st, c, k), work);
            }
            /* :588 xTOz keeps modes 0..nx, :590 FFT */
#pragma omp for collapse(2)
            for (int c = 0; c < 6; ++c)
                for (int ix = 0; ix <= nx; ++ix) {
                    cplx* line = VVDZ(st, slot, c, ix);
                    for (int k = 0; k < nzd; ++k) line[k] = VVDX(st, c, k)[ix];
                    fft_line(&st->pz, line, work, -1);
                }
        }
    }
    if (cflmax > st->cfl) st->cfl = cflmax;
}

static inline int imod(int iy) { return (iy + 1000) % 5; }

static inline cplx Dst(const co_state* st, const double* tab, int iy, const cplx* f, int c, int ix, int izp) {
    /* D0..D4 macros dnsdata.f90:330-333: real and imaginary parts separately */
    const cplx* col = f + IDXV(st, c, ix, izp, iy + 1 - 2);
    double re = 0, im = 0;
    for (int j = 0; j < 5; ++j) { re += tab[(iy + 1) * 5 + j] * creal(col[j]); im += tab[(iy + 1) * 5 + j] * cimag(col[j]); }
    return re + I * im;
}

void co_set_coriolis(co_state* st, double omega2, double kz_cutoff, double y_threshold_bot) {
    /* config_body_force, body_forces/coriolis/coriolis.inc:4-27 */
    st->bodyforce = 1;
    st->omega2 = omega2;
    st->y_thr_bot = y_threshold_bot;
    st->y_thr_top = st->ymax - y_threshold_bot;
    int thr = (int)floor(kz_cutoff / st->beta0);
    st->iz_thr = thr < st->nz ? thr : st->nz;
    if (!st->F) st->F = (cplx*)calloc((size_t)3 * (st->nx + 1) * st->nzt * st->nyp, sizeof(cplx));
}
/* config_body_force of body_forces/am_f1/am_f1.inc:4-10 and am_butterfly/am_butterfly.inc:4-9; which = 2 | 3 */
void co_set_am(co_state* st, int which, double lambdaz_f, double amp) {
    st->bodyforce = which;
    st->am_amp = amp;
    st->am_iz_f = (int)lround((2.0 * M_PI / lambdaz_f) / (st->beta0 / 1000.0));
    if (!st->F) st->F = (cplx*)calloc((size_t)3 * (st->nx + 1) * st->nzt * st->nyp, sizeof(cplx));
}
void co_set_body_force(co_state* st) {
    if (!st->bodyforce) return;
    if (st->bodyforce == 1) { /* coriolis.inc:29-41 */
#pragma omp parallel for
        for (int ix = 0; ix <= st->nx; ++ix)
            for (int iz = -st->iz_thr; iz <= st->iz_thr; ++iz)
                for (int iy = -1; iy <= st->ny + 1; ++iy) {
                    if (YY(st, iy) <= st->y_thr_bot || YY(st, iy) >= st->y_thr_top) {
                        for (int c = 1; c <= 2; ++c) {
                            const int pc = c % 2 + 1;
                            const int zeichen = 2 * pc - 3;
                            st->F[IDXV(st, pc - 1, ix, iz + st->nz, iy + 1)] = zeichen * st->omega2 * st->V[IDXV(st, c - 1, ix, iz + st->nz, iy + 1)];
                        }
                    }
                }
        return;
    }
    /* am_f1.inc:13-28 (loop over |iz| <= iz_f, condition lambda_z+ > 2.3 (y+)^2) and am_butterfly.inc:11-29 (all iz,
     * two boxes); both F = -amp V, the mean mode excluded where the hooks exclude it */
    const int izlim = st->bodyforce == 2 ? (st->am_iz_f < st->nz ? st->am_iz_f : st->nz) : st->nz;
#pragma omp parallel for
    for (int ix = 0; ix <= st->nx; ++ix)
        for (int iV = 0; iV < 3; ++iV)
            for (int iz = -izlim; iz <= izlim; ++iz)
                for (int iy = -1; iy <= st->ny + 1; ++iy) {
                    const double y = YY(st, iy);
                    const double yp = (y > 1 ? st->ymax - y : y) * 1000;
                    int on;
                    if (st->bodyforce == 2) {
                        const double lzp = iz == 0 ? 1e10 : 2 * M_PI / (st->beta0 * abs(iz)) * 1000;
                        on = (lzp > 2.3 * yp * yp) && !(ix == 0 && iz == 0);
                    } else {
                        on = (abs(iz) <= st->am_iz_f && yp <= 60 && !(ix == 0 && iz == 0)) || (abs(iz) > st->am_iz_f && yp > 60);
                    }
                    if (on) st->F[IDXV(st, iV, ix, iz + st->nz, iy + 1)] = -st->am_amp * st->V[IDXV(st, iV, ix, iz + st->nz, iy + 1)];
                }
}

/* dnsdata.f90:611-673.  max_iters < 0: the whole plane loop; otherwise stop after that many
 * iterations of it (timing sample; the state is then only partially advanced). */
static void buildrhs_n(co_state* st, const double* ODE, int compute_cfl, int max_iters) {
    const int nx = st->nx, ny = st->ny, nz = st->nz, nzt = st->nzt;
    const double deltat = st->deltat, ni = st->ni;
    if (st->bodyforce) { /* :616-629 */
#pragma omp parallel for collapse(2)
        for (int ix = 0; ix <= nx; ++ix)
            for (int izp = 0; izp < nzt; ++izp)
                for (int c = 0; c < 3; ++c) {
                    cplx* f = st->F + IDXV(st, c, ix, izp, 0);
                    f[0] = 0; f[1] = 0;
                    f[0] = -Dst(st, st->d4, 1, st->F, c, ix, izp) / DER(st->d4, 1, -2);
                    f[ny + 1] = 0; f[ny + 2] = 0;
                    f[ny + 2] = -Dst(st, st->d4, ny - 1, st->F, c, ix, izp) / DER(st->d4, ny - 1, 2);
                }
    }
    int iters = 0;
    for (int iy = -3; iy <= ny + 1; ++iy) {
        if (max_iters >= 0 && iters >= max_iters) break;
        ++iters;
        if (iy <= ny - 1) {
            double t0 = omp_get_wtime();
            convolutions(st, iy + 2, imod(iy + 2), compute_cfl, 1);
            double t1 = omp_get_wtime();
            st->t_conv += t1 - t0;
            if (iy >= 1) {
                const int sl[5] = {imod(iy - 2), imod(iy - 1), imod(iy), imod(iy + 1), imod(iy + 2)};
                const double* d0 = st->d0 + (iy + 1) * 5; const double* d1 = st->d1 + (iy + 1) * 5;
                const double* d2 = st->d2 + (iy + 1) * 5; const double* d4 = st->d4 + (iy + 1) * 5;
                const int nslot = (iy + 1000) % 3;
#pragma omp parallel for
                for (int ix = 0; ix <= nx; ++ix) {
                    const cplx ia = I * (ix * st->alfa0);
                    for (int iz = -nz; iz <= nz; ++iz) {
                        const int izp = iz + nz;
                        const cplx ib = I * (iz * st->beta0);
                        const double k2 = pow(st->alfa0 * ix, 2.0) + pow(st->beta0 * iz, 2.0);
                        const int kz = st->izd[izp];
#define DD(tab, k) (tab[0] * VVDZ(st, sl[0], (k) - 1, ix)[kz] + tab[1] * VVDZ(st, sl[1], (k) - 1, ix)[kz] + tab[2] * VVDZ(st, sl[2], (k) - 1, ix)[kz] + \
                    tab[3] * VVDZ(st, sl[3], (k) - 1, ix)[kz] + tab[4] * VVDZ(st, sl[4], (k) - 1, ix)[kz])
                        const cplx DD0_6 = DD(d0, 6), DD1_6 = DD(d1, 6);
                        const cplx rhsu = -ia * DD(d0, 1) - DD(d1, 4) - ib * DD0_6;
                        const cplx rhsv = -ia * DD(d0, 4) - DD(d1, 2) - ib * DD(d0, 5);
                        const cplx rhsw = -ia * DD0_6 - DD(d1, 5) - ib * DD(d0, 3);
                        cplx expl = ia * (ia * DD(d1, 1) + DD(d2, 4) + ib * DD1_6) + ib * (ia * DD1_6 + DD(d2, 5) + ib * DD(d1, 3)) - k2 * rhsv;
#undef DD
                        if (st->bodyforce)
                            expl = expl - k2 * Dst(st, st->d0, iy, st->F, 1, ix, izp) - ia * Dst(st, st->d1, iy, st->F, 0, ix, izp) - ib * Dst(st, st->d1, iy, st->F, 2, ix, izp);
                        cplx* newr = st->memrhs + (((size_t)ix * nzt + izp) * 3 + nslot) * 2;
                        cplx* oldr = st->oldrhs + (((size_t)ix * nzt + izp) * (ny - 1) + (iy - 1)) * 2;
                        const cplx* v = st->V + IDXV(st, 1, ix, izp, iy + 1 - 2);
                        const cplx* u = st->V + IDXV(st, 0, ix, izp, iy + 1 - 2);
                        const cplx* w = st->V + IDXV(st, 2, ix, izp, iy + 1 - 2);
                        { /* D2v, :647 */
                            const cplx unkn = Dst(st, st->d2, iy, st->V, 1, ix, izp) - k2 * Dst(st, st->d0, iy, st->V, 1, ix, izp);
                            cplx impl = 0;
                            for (int j = 0; j < 5; ++j) impl += (ni * (d4[j] - 2.0 * k2 * d2[j] + k2 * k2 * d0[j])) * v[j];
                            newr[1] = ODE[0] * unkn / deltat + impl + ODE[1] * expl - ODE[2] * oldr[1];
                            oldr[1] = expl;
                        }
                        if (ix == 0 && iz == 0) { /* :648-654 */
                            expl = (creal(rhsu) + st->meanpx) + I * (creal(rhsw) + st->meanpz);
                            double a0 = 0, b0 = 0, a2 = 0, b2 = 0;
                            for (int j = 0; j < 5; ++j) { a0 += d0[j] * creal(u[j]); b0 += d0[j] * creal(w[j]); a2 += d2[j] * creal(u[j]); b2 += d2[j] * creal(w[j]); }
                            if (st->bodyforce) {
                                const cplx* f1 = st->F + IDXV(st, 0, ix, izp, iy + 1 - 2); const cplx* f3 = st->F + IDXV(st, 2, ix, izp, iy + 1 - 2);
                                double fa = 0, fb = 0;
                                for (int j = 0; j < 5; ++j) { fa += d0[j] * creal(f1[j]); fb += d0[j] * creal(f3[j]); }
                                expl = expl + (fa + I * fb);
                            }
                            newr[0] = ODE[0] * (a0 + I * b0) / deltat + ni * (a2 + I * b2) + ODE[1] * expl - ODE[2] * oldr[0];
                            oldr[0] = expl;
                        } else { /* :656-661 */
                            expl = ib * rhsu - ia * rhsw;
                            if (st->bodyforce) expl = expl + ib * Dst(st, st->d0, iy, st->F, 0, ix, izp) - ia * Dst(st, st->d0, iy, st->F, 2, ix, izp);
                            const cplx unkn = ib * Dst(st, st->d0, iy, st->V, 0, ix, izp) - ia * Dst(st, st->d0, iy, st->V, 2, ix, izp);
                            cplx impl = 0;
                            for (int j = 0; j < 5; ++j) impl += (ni * (d2[j] - k2 * d0[j])) * (ib * u[j] - ia * w[j]);
                            newr[0] = ODE[0] * unkn / deltat + impl + ODE[1] * expl - ODE[2] * oldr[0];
                            oldr[0] = expl;
                        }
                    }
                }
                st->t_rhs += omp_get_wtime() - t1;
            }
        }
        if (iy - 2 >= 1) { /* :667-671 */
            double t1 = omp_get_wtime();
            const int nslot = (iy - 2 + 1000) % 3;
#pragma omp parallel for
            for (int ix = 0; ix <= nx; ++ix)
                for (int izp = 0; izp < nzt; ++izp) {
                    const cplx* nr = st->memrhs + (((size_t)ix * nzt + izp) * 3 + nslot) * 2;
                    st->V[IDXV(st, 0, ix, izp, iy - 2 + 1)] = nr[0];
                    st->V[IDXV(st, 1, ix, izp, iy - 2 + 1)] = nr[1];
                }
            st->t_rhs += omp_get_wtime() - t1;
        }
    }
}
void co_buildrhs(co_state* st, const double* ODE, int compute_cfl) { buildrhs_n(st, ODE, compute_cfl, -1); }

/* linsolve_blocking.inc:3-107 for ix in [0, nix) */
static void linsolve_n(co_state* st, double lambda, int nix) {
    const int nx = st->nx, ny = st->ny, nz = st->nz, nzt = st->nzt, nyp = st->nyp;
    const double ni = st->ni;
    (void)nx;
#pragma omp parallel
    {
        double* D2vmat = (double*)malloc(sizeof(double) * 5 * (ny + 1));
        double* etamat = (double*)malloc(sizeof(double) * 5 * (ny + 1));
        cplx* temp = (cplx*)malloc(sizeof(cplx) * nyp);
        cplx* ucor = (cplx*)malloc(sizeof(cplx) * nyp);
#pragma omp for collapse(2) schedule(static)
        for (int ix = 0; ix < nix; ++ix)
            for (int izp = 0; izp < nzt; ++izp) {
                const int iz = izp - nz;
                const cplx ia = I * (ix * st->alfa0), ib = I * (iz * st->beta0);
                const double k2 = pow(st->alfa0 * ix, 2.0) + pow(st->beta0 * iz, 2.0);
                cplx* v = st->V + IDXV(st, 1, ix, izp, 0);   /* index iy+1 */
                cplx* eta = st->V + IDXV(st, 0, ix, izp, 0);
                cplx* w3 = st->V + IDXV(st, 2, ix, izp, 0);
                memset(D2vmat + 5 * (ny - 1), 0, sizeof(double) * 10); /* halo rows ny, ny+1: zero (A.7) */
                memset(etamat + 5 * (ny - 1), 0, sizeof(double) * 10);
                for (int iy = 1; iy <= ny - 1; ++iy)
                    for (int j = 0; j < 5; ++j) {
                        const double d0 = st->d0[(iy + 1) * 5 + j], d2 = st->d2[(iy + 1) * 5 + j], d4 = st->d4[(iy + 1) * 5 + j];
                        const double OS = ni * (d4 - 2.0 * k2 * d2 + k2 * k2 * d0), SQ = ni * (d2 - k2 * d0);
                        D2vmat[(iy - 1) * 5 + j] = lambda * (d2 - k2 * d0) - OS; /* :12 */
                        etamat[(iy - 1) * 5 + j] = lambda * d0 - SQ;             /* :13 */
                    }
                /* wall data: bc0/bcn%u,w are zero except (0,0)%u = u0/uN (channel.f90:122-124, A.7) */
                cplx bc0_v = 0, bc0_vy = 0, bc0_eta = 0, bcn_v = 0, bcn_vy = 0, bcn_eta = 0;
                if (ix == 0 && iz == 0) { bc0_eta = st->u0; bcn_eta = st->uN; }
                bc0_v = bc0_v - st->v0bc[0] * bc0_vy / st->v0m1bc[0];                                    /* :21 */
                applybc_0(D2vmat, st->v0bc, st->v0m1bc);                                                 /* :22 */
                v[1 + 1] = v[1 + 1] - D2vmat[0] * bc0_vy / st->v0m1bc[0] - D2vmat[1] * bc0_v / st->v0bc[1]; /* :23 */
                v[2 + 1] = v[2 + 1] - D2vmat[5 + 0] * bc0_v / st->v0bc[1];                               /* :24 */
                applybc_0(etamat, st->eta0bc, st->eta0m1bc);                                             /* :25 */
                eta[1 + 1] = eta[1 + 1] - etamat[1] * bc0_eta / st->eta0bc[1];                           /* :26 */
                eta[2 + 1] = eta[2 + 1] - etamat[5 + 0] * bc0_eta / st->eta0bc[1];                       /* :27 */
                bcn_v = bcn_v - st->vnbc[4] * bcn_vy / st->vnp1bc[4];                                    /* :35 */
                applybc_n(D2vmat, ny, st->vnbc, st->vnp1bc);                                             /* :36 */
                v[ny - 1 + 1] = v[ny - 1 + 1] - D2vmat[(ny - 2) * 5 + 4] * bcn_vy / st->vnp1bc[4] - D2vmat[(ny - 2) * 5 + 3] * bcn_v / st->vnbc[3]; /* :37 */
                v[ny - 2 + 1] = v[ny - 2 + 1] - D2vmat[(ny - 3) * 5 + 4] * bcn_v / st->vnbc[3];          /* :38 */
                applybc_n(etamat, ny, st->etanbc, st->etanp1bc);                                         /* :39 */
                eta[ny - 1 + 1] = eta[ny - 1 + 1] - etamat[(ny - 2) * 5 + 3] * bcn_eta / st->etanbc[3];  /* :40 */
                eta[ny - 2 + 1] = eta[ny - 2 + 1] - etamat[(ny - 3) * 5 + 4] * bcn_eta / st->etanbc[3];  /* :41 */
                LU5decompStep(D2vmat, ny + 1); LU5decompStep(etamat, ny + 1);                            /* :43 */
                LeftLU5divStep1(v, D2vmat, ny + 1);                                                      /* :44 */
                LeftLU5divStep1(eta, etamat, ny + 1);                                                    /* :45 */
                LeftLU5divStep2(D2vmat, v, ny + 1);                                                      /* :48 */
                LeftLU5divStep2(etamat, eta, ny + 1);                                                    /* :49 */
#define S3(f, i0, bc, o) ((f)[(i0) + 1] * (bc)[o] + (f)[(i0) + 2] * (bc)[(o) + 1] + (f)[(i0) + 3] * (bc)[(o) + 2])
#define S4(f, i0, bc, o) (S3(f, i0, bc, o) + (f)[(i0) + 4] * (bc)[(o) + 3])
                v[0 + 1] = (bc0_v - S3(v, 1, st->v0bc, 2)) / st->v0bc[1];                                /* :51 */
                v[-1 + 1] = (bc0_vy - S4(v, 0, st->v0m1bc, 1)) / st->v0m1bc[0];                          /* :52 */
                eta[0 + 1] = (bc0_eta - S3(eta, 1, st->eta0bc, 2)) / st->eta0bc[1];                      /* :53 */
                eta[-1 + 1] = -S4(eta, 0, st->eta0m1bc, 1) / st->eta0m1bc[0];                            /* :54 */
                v[ny + 1] = (bcn_v - S3(v, ny - 3, st->vnbc, 0)) / st->vnbc[3];                          /* :57 */
                v[ny + 1 + 1] = (bcn_vy - S4(v, ny - 3, st->vnp1bc, 0)) / st->vnp1bc[4];                 /* :58 */
                eta[ny + 1] = (bcn_eta - S3(eta, ny - 3, st->etanbc, 0)) / st->etanbc[3];                /* :59 */
                eta[ny + 1 + 1] = -S4(eta, ny - 3, st->etanp1bc, 0) / st->etanp1bc[4];                   /* :60 */
                if (ix == 0 && iz == 0) { /* :62-97 */
                    for (int i = 0; i < nyp; ++i) { w3[i] = cimag(eta[i]); eta[i] = creal(eta[i]); }
                    for (int i = 0; i < nyp; ++i) ucor[i] = 0;
                    for (int iy = 1; iy <= ny - 1; ++iy) ucor[iy + 1] = 1;
                    LeftLU5divStep1(ucor, etamat, ny + 1);
                    LeftLU5divStep2(etamat, ucor, ny + 1);
                    ucor[0 + 1] = -S3(ucor, 1, st->eta0bc, 2) / st->eta0bc[1];
                    ucor[-1 + 1] = -S4(ucor, 0, st->eta0m1bc, 1) / st->eta0m1bc[0];
                    ucor[ny + 1] = -S3(ucor, ny - 3, st->etanbc, 0) / st->etanbc[3];
                    ucor[ny + 1 + 1] = -S4(ucor, ny - 3, st->etanp1bc, 0) / st->etanp1bc[4];
                    st->fr[0] = yintegr(st, (const double*)eta, 2);
                    st->fr[1] = yintegr(st, (const double*)w3, 2);
                    st->fr[2] = yintegr(st, (const double*)ucor, 2);
                    if (fabs(st->meanflowx) > 1.0e-7 && !st->CPI) {
                        st->corrpx = (st->meanflowx - st->fr[0]) / st->fr[2];
                        for (int i = 0; i < nyp; ++i) eta[i] = (creal(eta[i]) + st->corrpx * creal(ucor[i])) + I * cimag(eta[i]);
                    }
                    if (fabs(st->meanflowz) > 1.0e-7 && !st->CPI) {
                        st->corrpz = (st->meanflowz - st->fr[1]) / st->fr[2];
                        for (int i = 0; i < nyp; ++i) w3[i] = (creal(w3[i]) + st->corrpz * creal(ucor[i])) + I * cimag(w3[i]);
                    }
                    if (st->CPI) {
                        if (st->CPI_type == 0) st->meanpx = (1 - st->gamma) * 6 * ni / st->fr[0];
                        else if (st->CPI_type == 1) st->meanpx = (1.5 / st->gamma) * st->fr[0] * ni;
                    }
                } else { /* :99-103 with COMPLEXderiv dnsdata.f90:349-371 */
                    const cplx* f0 = v;
                    cplx* f1 = w3;
                    f1[0 + 1] = 0; f1[-1 + 1] = 0; f1[ny + 1] = 0; f1[ny + 2] = 0;
                    cplx a = 0, b = 0, c = 0, d = 0;
                    for (int j = 0; j < 5; ++j) {
                        a += st->d140[j] * f0[j]; b += st->d14m1[j] * f0[j];
                        c += st->d14n[j] * f0[ny - 3 + 1 + j]; d += st->d14np1[j] * f0[ny - 3 + 1 + j];
                    }
                    f1[0 + 1] = a; f1[-1 + 1] = b; f1[ny + 1] = c; f1[ny + 2] = d;
                    for (int iy = 1; iy <= ny - 1; ++iy) {
                        cplx s = 0;
                        for (int j = 0; j < 5; ++j) s += st->d1[(iy + 1) * 5 + j] * f0[iy - 2 + 1 + j];
                        f1[iy + 1] = s;
                    }
                    f1[1 + 1] = f1[1 + 1] - (DER(st->d0, 1, -1) * f1[0 + 1] + DER(st->d0, 1, -2) * f1[-1 + 1]);
                    f1[2 + 1] = f1[2 + 1] - DER(st->d0, 2, -2) * f1[0 + 1];
                    f1[ny - 1 + 1] = f1[ny - 1 + 1] - (DER(st->d0, ny - 1, 1) * f1[ny + 1] + DER(st->d0, ny - 1, 2) * f1[ny + 2]);
                    f1[ny - 2 + 1] = f1[ny - 2 + 1] - DER(st->d0, ny - 2, 2) * f1[ny + 1];
                    LeftLU5divStep1(f1, st->D0mat, ny + 1);
                    LeftLU5divStep2(st->D0mat, f1, ny + 1);
                    for (int i = 0; i < nyp; ++i) temp[i] = (ia * f1[i] - ib * eta[i]) / k2;
                    for (int i = 0; i < nyp; ++i) f1[i] = (ib * f1[i] + ia * eta[i]) / k2;
                    for (int i = 0; i < nyp; ++i) eta[i] = temp[i];
                }
            }
        free(D2vmat); free(etamat); free(temp); free(ucor);
    }
}
void co_linsolve(co_state* st, double lambda) { linsolve_n(st, lambda, st->nx + 1); }

static void flowrate_cpi(co_state* st) { /* channel.f90:101-115 */
    st->fr[0] = yintegr(st, (const double*)(st->V + IDXV(st, 0, 0, st->nz, 0)), 2);
    st->fr[1] = yintegr(st, (const double*)(st->V + IDXV(st, 2, 0, st->nz, 0)), 2);
    if (st->CPI) {
        if (st->CPI_type == 0) st->meanpx = (1 - st->gamma) * 6 * st->ni / st->fr[0];
        else if (st->CPI_type == 1) st->meanpx = (1.5 / st->gamma) * st->fr[0] * st->ni;
    }
}
void co_cfl_prepass(co_state* st) { /* channel.f90:95-115 */
    if (st->deltat == 0) st->deltat = 1.0;
    for (int iy = 1; iy <= st->ny - 1; ++iy) convolutions(st, iy, 0, 1, 0);
    flowrate_cpi(st);
}

void co_outstats(co_state* st, double* line) { /* dnsdata.f90:853-880 */
    const int ny = st->ny;
    const double rg = st->cfl;
    st->cfl = 0;
    if (st->cflmax > 0) st->deltat = st->cflmax / rg;
    const cplx* U = st->V + IDXV(st, 0, 0, st->nz, 0);
    const cplx* W = st->V + IDXV(st, 2, 0, st->nz, 0);
    double dudy0 = 0, dwdy0 = 0, dudyN = 0, dwdyN = 0;
    for (int j = 0; j < 5; ++j) {
        dudy0 += st->d140[j] * creal(U[j]); dwdy0 += st->d140[j] * creal(W[j]);
        dudyN += st->d14n[j] * creal(U[ny - 3 + 1 + j]); dwdyN += st->d14n[j] * creal(W[ny - 3 + 1 + j]);
    }
    line[0] = st->time; line[1] = dudy0; line[2] = -dudyN; line[3] = dwdy0; line[4] = -dwdyN;
    line[5] = st->fr[0] + st->corrpx * st->fr[2]; line[6] = st->meanpx + st->corrpx;
    line[7] = st->fr[1] + st->corrpz * st->fr[2]; line[8] = st->meanpz + st->corrpz;
    line[9] = rg * st->deltat; line[10] = st->deltat;
}

static const double RK_rai[3][3] = {{120.0 / 32.0, 2.0, 0.0}, {120.0 / 8.0, 50.0 / 8.0, 34.0 / 8.0}, {120.0 / 20.0, 90.0 / 20.0, 50.0 / 20.0}}; /* dnsdata.f90:70-72 */

void co_step(co_state* st, double* line) { /* channel.f90:118-167 */
    for (int k = 0; k < 3; ++k) {
        st->time = st->time + 2.0 / RK_rai[k][0] * st->deltat;
        if (st->bodyforce) co_set_body_force(st);
        co_buildrhs(st, RK_rai[k], k == 2);
        co_linsolve(st, RK_rai[k][0] / st->deltat);
    }
    co_outstats(st, line);
}

/* Timed, bounded sample of one RK substep (substep 2 coefficients): `iters` iterations of the
 * buildrhs plane loop and the columns of `nix` x-modes of linsolve.  Returns per-part seconds and
 * how many planes/columns each covered so the caller can scale to a full step.  The state is
 * left partially advanced: use a scratch state. */
void co_sample(co_state* st, int iters, int nix, double* out) {
    st->t_conv = st->t_rhs = st->t_solve = 0;
    if (st->deltat == 0) st->deltat = 1e-3;
    buildrhs_n(st, RK_rai[1], 0, iters);
    int nconv = 0, nrhs = 0;
    for (int i = 0, iy = -3; i < iters && iy <= st->ny + 1; ++i, ++iy) {
        if (iy <= st->ny - 1) { ++nconv; if (iy >= 1) ++nrhs; }
    }
    if (nix > st->nx + 1) nix = st->nx + 1;
    double t0 = omp_get_wtime();
    linsolve_n(st, RK_rai[1][0] / st->deltat, nix);
    st->t_solve = omp_get_wtime() - t0;
    out[0] = st->t_conv; out[1] = nconv; out[2] = st->t_rhs; out[3] = nrhs; out[4] = st->t_solve; out[5] = nix;
}

/* synthetic smooth field for the timing sample (values are irrelevant to the timing, they only
 * need to be finite): u = laminar + small modes */
void co_fill_synthetic(co_state* st) {
    const int nx = st->nx, nz = st->nz, nyp = st->nyp;
#pragma omp parallel for collapse(2)
    for (int c = 0; c < 3; ++c)
        for (int ix = 0; ix <= nx; ++ix)
            for (int izp = 0; izp < st->nzt; ++izp) {
                const double k2 = pow(st->alfa0 * ix, 2.0) + pow(st->beta0 * (izp - nz), 2.0);
                const double amp = 1e-3 / (1.0 + k2);
                unsigned h = (unsigned)(c * 73856093u) ^ (unsigned)(ix * 19349663u) ^ (unsigned)(izp * 83492791u);
                const double p1 = (h & 1023) / 1024.0 - 0.5, p2 = ((h >> 10) & 1023) / 1024.0 - 0.5;
                cplx* col = st->V + IDXV(st, c, ix, izp, 0);
                for (int i = 0; i < nyp; ++i) {
                    const double yy = st->y[i], g = (yy * (2 - yy)) * (yy * (2 - yy));
                    col[i] = amp * g * (p1 + I * p2);
                }
            }
    cplx* U = st->V + IDXV(st, 0, 0, nz, 0);
    for (int i = 0; i < nyp; ++i) U[i] = 1.5 * st->y[i] * (2 - st->y[i]);
    cplx* Vm = st->V + IDXV(st, 1, 0, nz, 0);
    cplx* Wm = st->V + IDXV(st, 2, 0, nz, 0);
    for (int i = 0; i < nyp; ++i) { Vm[i] = 0; Wm[i] = 0; }
}
