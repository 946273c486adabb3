// fft_regs.cuh - register-resident FP64 complex FFT building blocks for sm_100a.
//
// Replaces the FFTW3 plans of the reference (ffts.f90:70-75) for the large transform sizes.
// A line of N = A*B*C points is transformed in three stages of register-resident radix-A/B/C
// DFTs; between stages the data is exchanged once through shared memory (two exchanges per
// transform, 64 B of shared-memory traffic per point instead of one round trip per radix-8
// pass), and the first/last stage reads/writes global memory directly.
//
//   decimation in frequency (natural in, scrambled registers out):
//     n = a*BC + b*C + c  ->  DFT_A over a, * w_N^(ka*(b*C+c))  ->  DFT_B over b, * w_BC^(kb*c)
//                         ->  DFT_C over c;   result index k = ka + A*kb + A*B*kc
//   decimation in time is the exact transpose (stages C, B, A with the twiddles before the
//   butterflies): scrambled registers in, natural order out.  Pointwise work in between (the
//   velocity products, dnsdata.f90:581-584) does not care about the order.
//
// Shared-memory layouts of one line (cplx units), both conflict-free for 16-byte accesses:
//   L1(ka, t1)      = ka*BC + t1                 t1 = b*C + c        (between stages A and B)
//   L2(ka, kb, c)   = (kb*A + ka)*(C+1) + c                          (between stages B and C)
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include "fft_device.cuh"

// ---- compile-time loops ---------------------------------------------------------------------
template <int I>
struct IC {
    static constexpr int value = I;
    __host__ __device__ constexpr operator int() const { return I; }
};
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(IC<Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- twiddle constants: cos/sin(2 pi e / R) for R | 48 (multiples of 7.5 degrees) -------------
__host__ __device__ constexpr double chb_cos48(int m) {  // cos(2 pi m / 48), m in 0..47
    m = ((m % 48) + 48) % 48;
    if (m > 24) m = 48 - m;            // cos symmetric
    bool neg = false;
    if (m > 12) { m = 24 - m; neg = true; }
    double v = 0.0;
    switch (m) {
        case 0: v = 1.0; break;
        case 1: v = 0.99144486137381041114455752692856; break;   //  7.5
        case 2: v = 0.96592582628906828674974319972890; break;   // 15
        case 3: v = 0.92387953251128675612818318939679; break;   // 22.5
        case 4: v = 0.86602540378443864676372317075294; break;   // 30
        case 5: v = 0.79335334029123516457977696150130; break;   // 37.5
        case 6: v = 0.70710678118654752440084436210485; break;   // 45
        case 7: v = 0.60876142900872063941609754289816; break;   // 52.5
        case 8: v = 0.5; break;                                  // 60
        case 9: v = 0.38268343236508977172845998403040; break;   // 67.5
        case 10: v = 0.25881904510252076234889883762405; break;  // 75
        case 11: v = 0.13052619222005159154840622789549; break;  // 82.5
        default: v = 0.0; break;                                 // 90
    }
    return neg ? -v : v;
}
__host__ __device__ constexpr double chb_sin48(int m) { return chb_cos48(m - 12); }

// a * exp(S * 2 pi i * E / R)
template <int R, int E, int S>
__device__ __forceinline__ cplx mulw(cplx a) {
    constexpr int e = ((E % R) + R) % R;
    if constexpr (e == 0) {
        return a;
    } else if constexpr (4 * e == R) {
        return crot<S>(a);
    } else if constexpr (2 * e == R) {
        return make_double2(-a.x, -a.y);
    } else if constexpr (4 * e == 3 * R) {
        return crot<-S>(a);
    } else {
        static_assert(48 % R == 0, "mulw: unsupported radix");
        constexpr double c = chb_cos48(e * (48 / R));
        constexpr double s = (S > 0 ? 1.0 : -1.0) * chb_sin48(e * (48 / R));
        return make_double2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

// ---- composite radices on top of Dft<2|3|4|8> (fft_device.cuh) -------------------------------
// R = P*Q:  n = Q*n1 + n2,  k = k1 + P*k2
template <int P, int Q, int S>
__device__ __forceinline__ void dft_composite(cplx* a) {
    constexpr int R = P * Q;
    cplx y[R];
    static_for<Q>([&](auto n2) {
        cplx t[P];
        static_for<P>([&](auto n1) { t[n1] = a[Q * n1 + n2]; });
        Dft<P, S>::run(t);
        static_for<P>([&](auto k1) { y[k1 * Q + n2] = mulw<R, n2 * k1, S>(t[k1]); });
    });
    static_for<P>([&](auto k1) {
        cplx u[Q];
        static_for<Q>([&](auto n2) { u[n2] = y[k1 * Q + n2]; });
        Dft<Q, S>::run(u);
        static_for<Q>([&](auto k2) { a[k1 + P * k2] = u[k2]; });
    });
}
template <int S>
struct Dft<6, S> {
    __device__ __forceinline__ static void run(cplx* a) { dft_composite<2, 3, S>(a); }
};
// 12 = 3 x 4 with coprime factors: Good-Thomas prime-factor mapping, no twiddles between the 3- and
// 4-point transforms.  n = (4 n1 + 3 n2) mod 12, k = (4 k1 + 9 k2) mod 12:
//   n k = 4 n1 k1 + 3 n2 k2 (mod 12)  ->  X[k] = sum_{n2} w4^(n2 k2) sum_{n1} w3^(n1 k1) x[n]
template <int S>
struct Dft<12, S> {
    __device__ __forceinline__ static void run(cplx* a) {
        cplx y[12];   // y[k1*4 + n2]
        static_for<4>([&](auto n2_) {
            constexpr int n2 = decltype(n2_)::value;
            cplx t[3] = {a[(3 * n2) % 12], a[(4 + 3 * n2) % 12], a[(8 + 3 * n2) % 12]};
            Dft<3, S>::run(t);
            y[0 * 4 + n2] = t[0];
            y[1 * 4 + n2] = t[1];
            y[2 * 4 + n2] = t[2];
        });
        static_for<3>([&](auto k1_) {
            constexpr int k1 = decltype(k1_)::value;
            cplx u[4] = {y[k1 * 4 + 0], y[k1 * 4 + 1], y[k1 * 4 + 2], y[k1 * 4 + 3]};
            Dft<4, S>::run(u);
            static_for<4>([&](auto k2_) {
                constexpr int k2 = decltype(k2_)::value;
                a[(4 * k1 + 9 * k2) % 12] = u[k2];
            });
        });
    }
};
template <int S>
struct Dft<16, S> {
    __device__ __forceinline__ static void run(cplx* a) { dft_composite<4, 4, S>(a); }
};

// w[p] = w1^p for p = 1..R-1 (w[0] unused), shallow product tree
template <int R>
__device__ __forceinline__ void twiddle_powers(cplx w1, cplx* w) {
    w[1] = w1;
    static_for<R - 2>([&](auto i) {
        constexpr int p = i + 2;
        w[p] = cmul(w[p / 2], w[p - p / 2]);
    });
}

// ---- three-stage geometry -------------------------------------------------------------------
template <int N_, int A_, int B_, int C_>
struct Fft3 {
    static constexpr int N = N_, A = A_, B = B_, C = C_;
    static constexpr int BC = B * C, AB = A * B, CP = C + 1;
    static constexpr int NBF_A = N / A, NBF_B = N / B, NBF_C = N / C;  // butterflies per stage
    static constexpr int LINE = (N / C) * CP;                           // cplx per line buffer (>= N)
    static_assert(A * B * C == N, "bad factorisation");
    __device__ __forceinline__ static int L1(int ka, int t1) { return ka * BC + t1; }
    __device__ __forceinline__ static int L2(int ka, int kb, int c) { return (kb * A + ka) * CP + c; }
};

// Stage A (DIF): x[a], a = 0..A-1 are the inputs n = a*BC + t1 of butterfly t1; on return x[ka]
// is twiddled and ready for L1(ka, t1).  W = exp(+2 pi i e / N) table.
template <class G, int S>
__device__ __forceinline__ void dif_stage_a(cplx* x, int t1, const cplx* __restrict__ W) {
    Dft<G::A, S>::run(x);
    if (t1 != 0) {
        cplx w[G::A];
        twiddle_powers<G::A>(ctw<S>(W, t1), w);
        static_for<G::A - 1>([&](auto i) { x[i + 1] = cmul(x[i + 1], w[i + 1]); });
    }
}

// Stage B (DIF): x[b] are the values L1(ka, b*C + c) of butterfly (ka, c); wc[kb] = w_BC^(kb*c)
// (sign applied) for this thread's c.
template <class G, int S>
__device__ __forceinline__ void dif_stage_b(cplx* x, const cplx* wc, bool c_nonzero) {
    Dft<G::B, S>::run(x);
    if (c_nonzero) static_for<G::B - 1>([&](auto i) { x[i + 1] = cmul(x[i + 1], wc[i + 1]); });
}

#ifdef CHB_HOST_EMUL
// tests/host_emul: the asynchronous copies become plain copies done by the issuing thread (every use is followed by
// a __syncthreads() before other threads read the data), the mbarrier becomes a no-op.
#include <cstring>
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_prefetch_l2(const void*, unsigned) {}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
__device__ __forceinline__ void cp_async_wait_all() {}
#else
// ---- TMA 1-D bulk copy global -> shared with mbarrier completion (cp.async.bulk, SASS UBLKCP) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// L2 prefetch of a contiguous range (cp.async.bulk.prefetch.L2, SASS UBLKPF): one instruction, no
// destination; used to pull the tile of a CTA that will run a few hundred CTAs later into L2 so that
// HBM keeps streaming while the resident CTAs are in their compute phases.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- per-thread 16-byte asynchronous copies global -> shared (cp.async, SASS LDGSTS) ------------
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif   // CHB_HOST_EMUL
