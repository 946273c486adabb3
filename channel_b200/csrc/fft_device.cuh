// fft_device.cuh - block-cooperative FP64 complex FFTs on lines held in shared memory.
//
// Replaces the FFTW3 plans of the reference (ffts.f90:70-75).  Sizes are 2^a*3^b,
// b<=1 (fftFIT, ffts.f90:78-86).  Transforms are in place and unnormalised:
//   DIF (decimation in frequency): natural-order input  -> digit-reversed output
//   DIT (decimation in time)     : digit-reversed input -> natural-order output
// Pointwise work between an inverse and a forward transform (the velocity products,
// dnsdata.f90:581-584) is order independent, so the x-pass runs DIF -> products -> DIT
// and never permutes; the z-passes absorb the permutation into their global stores.
//
// Digit reversal for radices r_0..r_{k-1} (DIF pass order):
//   k = p_0 + r_0*(p_1 + r_1*(p_2 + ...))   <->   pos = sum_t p_t * N/(r_0*...*r_t)
#pragma once
#include <cuda_runtime.h>

typedef double2 cplx;

// dynamic shared memory of a kernel; tests/host_emul/cta_emul.hpp supplies the buffer when the kernel
// source is compiled for the CPU
#ifdef CHB_HOST_EMUL
#define CHB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cta_emul::g_dyn_smem)
#else
#define CHB_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

#define CHB_MAX_PASSES 10
struct FftPlan {
    int n;
    int npass;
    int radix[CHB_MAX_PASSES];
};

// one 16-byte element of padding every 8 elements keeps the small-stride passes
// (stride R elements between neighbouring threads) free of shared-memory bank conflicts
#define CHB_PAD(e) ((e) + ((e) >> 3))
static inline int chb_padded_len(int n) { return n + (n >> 3) + 1; }

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
// multiply by S*i
template <int S>
__device__ __forceinline__ cplx crot(cplx a) {
    return (S > 0) ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// table holds exp(+2 pi i e / N); S<0 conjugates
template <int S>
__device__ __forceinline__ cplx ctw(const cplx* __restrict__ W, int e) {
    cplx w = __ldg(&W[e]);
    if (S < 0) w.y = -w.y;
    return w;
}

template <int R, int S>
struct Dft;

template <int S>
struct Dft<2, S> {
    __device__ __forceinline__ static void run(cplx* a) {
        cplx t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};
template <int S>
struct Dft<3, S> {
    __device__ __forceinline__ static void run(cplx* a) {
        const double s3 = 0.86602540378443864676372317075294;  // sqrt(3)/2
        cplx t = cadd(a[1], a[2]);
        cplx d = csub(a[1], a[2]);
        cplx m = make_double2(a[0].x - 0.5 * t.x, a[0].y - 0.5 * t.y);
        cplx r = crot<S>(cscale(d, s3));
        a[0] = cadd(a[0], t);
        a[1] = cadd(m, r);
        a[2] = csub(m, r);
    }
};
template <int S>
struct Dft<4, S> {
    __device__ __forceinline__ static void run(cplx* a) {
        cplx t0 = cadd(a[0], a[2]);
        cplx t1 = csub(a[0], a[2]);
        cplx t2 = cadd(a[1], a[3]);
        cplx t3 = crot<S>(csub(a[1], a[3]));
        a[0] = cadd(t0, t2);
        a[1] = cadd(t1, t3);
        a[2] = csub(t0, t2);
        a[3] = csub(t1, t3);
    }
};
template <int S>
struct Dft<8, S> {
    __device__ __forceinline__ static void run(cplx* a) {
        const double h = 0.70710678118654752440084436210485;  // 1/sqrt(2)
        cplx e[4] = {a[0], a[2], a[4], a[6]};
        cplx o[4] = {a[1], a[3], a[5], a[7]};
        Dft<4, S>::run(e);
        Dft<4, S>::run(o);
        // o[p] *= w8^p, w8 = (1 + S i)/sqrt(2)
        cplx r1 = crot<S>(o[1]);
        o[1] = make_double2((o[1].x + r1.x) * h, (o[1].y + r1.y) * h);
        o[2] = crot<S>(o[2]);
        cplx r3 = crot<S>(o[3]);
        o[3] = make_double2((r3.x - o[3].x) * h, (r3.y - o[3].y) * h);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            a[p] = cadd(e[p], o[p]);
            a[p + 4] = csub(e[p], o[p]);
        }
    }
};

// One in-place radix-R pass over `nlines` lines (line l at buf + l*line_stride, padded
// indexing).  n_sub = size of the sub-transforms this pass splits/combines.
template <int R, int S, bool DIF>
__device__ __forceinline__ void fft_pass(cplx* __restrict__ buf, int line_stride, int nlines, int N, int n_sub,
                                         const cplx* __restrict__ W) {
    const int m = n_sub / R;
    const int nb = N / R;
    const int tws = N / n_sub;
    const int total = nlines * nb;
    for (int g = threadIdx.x; g < total; g += blockDim.x) {
        const int line = g / nb;
        const int gg = g - line * nb;
        const int b = gg / m;
        const int j = gg - b * m;
        cplx* x = buf + (size_t)line * line_stride;
        const int base = b * n_sub + j;
        cplx a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = x[CHB_PAD(base + q * m)];
        if (!DIF && j != 0) {
#pragma unroll
            for (int q = 1; q < R; ++q) a[q] = cmul(a[q], ctw<S>(W, q * j * tws));
        }
        Dft<R, S>::run(a);
        if (DIF && j != 0) {
#pragma unroll
            for (int p = 1; p < R; ++p) a[p] = cmul(a[p], ctw<S>(W, p * j * tws));
        }
#pragma unroll
        for (int p = 0; p < R; ++p) x[CHB_PAD(base + p * m)] = a[p];
    }
}

template <int S, bool DIF>
__device__ __forceinline__ void fft_pass_dispatch(int R, cplx* buf, int line_stride, int nlines, int N, int n_sub,
                                                  const cplx* W) {
    switch (R) {
        case 8: fft_pass<8, S, DIF>(buf, line_stride, nlines, N, n_sub, W); break;
        case 4: fft_pass<4, S, DIF>(buf, line_stride, nlines, N, n_sub, W); break;
        case 3: fft_pass<3, S, DIF>(buf, line_stride, nlines, N, n_sub, W); break;
        default: fft_pass<2, S, DIF>(buf, line_stride, nlines, N, n_sub, W); break;
    }
}

// In-place transform of `nlines` lines.  All threads of the block must call it; the
// data must be visible (a __syncthreads() precedes the first pass here) and is
// visible to all threads on return.
template <int S, bool DIF>
__device__ __forceinline__ void fft_lines(cplx* buf, int line_stride, int nlines, const FftPlan& pl,
                                          const cplx* __restrict__ W) {
    __syncthreads();
    if (DIF) {
        int n_sub = pl.n;
        for (int t = 0; t < pl.npass; ++t) {
            const int R = pl.radix[t];
            fft_pass_dispatch<S, true>(R, buf, line_stride, nlines, pl.n, n_sub, W);
            __syncthreads();
            n_sub /= R;
        }
    } else {
        int n_sub = 1;
        for (int t = pl.npass - 1; t >= 0; --t) {
            const int R = pl.radix[t];
            n_sub *= R;
            fft_pass_dispatch<S, false>(R, buf, line_stride, nlines, pl.n, n_sub, W);
            __syncthreads();
        }
    }
}
