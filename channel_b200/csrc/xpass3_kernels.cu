// xpass3_kernels.cu - x-direction pass of the nonlinear term for the large transform sizes:
//   zero-pad in x + c2r (RFT ffts.f90:72, dnsdata.f90:535) -> CFL (dnsdata.f90:552-556) -> the six
//   products * factor (dnsdata.f90:581-584) -> r2c (HFT ffts.f90:74) -> keep modes 0..nx
//   (xTOz pack, mpi_transpose.f90:99-106)
// for one physical z-line per group of T = nxd/C threads, fused in one kernel: physical-space data
// exists only in registers.
//
// A real transform of logical length 2M (M = nxd) is a complex transform of length M plus a
// split/merge pass:
//   c2r:  Z[n] = (X[n] + conj X[M-n]) + i e^{+i pi n/M} (X[n] - conj X[M-n]),  z = DFT+_M(Z),
//         r[2m] = Re z[m], r[2m+1] = Im z[m]          (imaginary parts of X[0], X[M] ignored; X[M]=0)
//   r2c:  Z = DFT-_M(z),  X[j] = (Z[j] + conj Z[M-j])/2 - (i/2) e^{-i pi j/M} (Z[j] - conj Z[M-j])
// The backward transform runs decimation-in-frequency (natural in, scrambled registers out), the
// forward one is its transpose (scrambled in, natural out), so the pointwise products sit between
// the two innermost radix-C butterflies without any reordering (fft_regs.cuh).
//
// Shared memory per line: six buffers of A*(BC+1) complex (u,v,w, later the six products, in
// place).  Per point and transform: two exchanges (backward) / two exchanges + the merge pass
// (forward) through shared memory; global memory is touched once per input and output mode.
#include <cstdlib>
#include "chb_internal.h"
#include "fft_regs.cuh"

// x[p] *= w1^p, p = 1..R-1, sequential recurrence (low register pressure)
template <int R>
__device__ __forceinline__ void apply_twiddle_seq(cplx* x, cplx w1) {
    cplx wp = w1;
    x[1] = cmul(x[1], wp);
    static_for<R - 2>([&](auto i_) {
        constexpr int p = decltype(i_)::value + 2;
        wp = cmul(wp, w1);
        x[p] = cmul(x[p], wp);
    });
}

// ---------------------------------------------------------------------------------------------
// Every stage's butterflies are flattened over (component, butterfly) tasks, so that a small innermost
// radix C (T = M/C threads per line, e.g. C = 4 -> 192 threads for M = 768) keeps all threads busy in
// every stage and MINB CTAs (lines) per SM give 4-5 warps per scheduler.
//
// SPLIT = 2 (CHB_XPASS_SPLIT; default at nxd = 1536, where it measured 6 % faster; 34 % slower at nxd = 768): TWO threads per
// innermost butterfly position (T = 2 M/C threads per line).  For M = 1536, whose six line buffers (148.6 KB) leave
// room for one CTA per SM only, that is twice the resident warps at the same shared memory (12 instead of 6, 140
// registers); for M = 768, two CTAs of 12 warps at 80 registers instead of three of 6 warps at 96.  The flattened
// task loops of the outer stages take the extra threads as they are; in the innermost stage both threads of a
// position run the radix-C butterflies of u, v, w (duplicated: 3 of the 9 innermost butterflies per position), then
// one forms uu, vv, ww and the other uv, vw, uw.  Because two threads now read the u, v, w entries that the products
// overwrite in place, a barrier separates the reads from the writes.
// PERSIST (CHB_XPASS_PERSIST, default at nxd = 1536): a persistent CTA walks over the lines of the launch.  The
// inputs of line n+1 (3 x (nx+1) modes) are copied into a shared-memory staging area with per-thread 16-byte
// cp.async while line n runs its stages B ... merge, and the twiddle tables (Wh[0..nx], W[0..BC), W[A k]) are loaded
// into shared memory once per CTA, so that no stage waits for a global load any more: at nxd = 1536, where the six
// line buffers (148.6 KB) leave one CTA of 6-12 warps per SM, 23 % of the warp samples were long-scoreboard stalls of
// stage A and of the merge pass (profiles/r2b_ncu_source_xpass.md).  216 KB of shared memory at nxd = 1536.
template <class G, int LPC, int MINB, bool MULTI, int SPLIT = 1, bool PERSIST = false>
__global__ void __launch_bounds__(SPLIT * LPC * (G::N / G::C), MINB)
xpass4_kernel(const cplx* __restrict__ Ar, const __grid_constant__ PeerPtrs Bw, const __grid_constant__ Geometry g, const cplx* __restrict__ W,
              const cplx* __restrict__ Wh, const double* __restrict__ dy, DevScalars* sc, int plane0, int np,
              int compute_cfl, int nplanes) {
    constexpr int M = G::N, A = G::A, B = G::B, C = G::C, BC = G::BC;
    constexpr int TP = M / C;          // butterfly positions of the innermost stage
    constexpr int T = SPLIT * TP;      // threads per line
    constexpr int BCP = BC + 1;
    constexpr int LB = A * BCP;        // complex per buffer
    constexpr int NB = A * C;          // stage-B butterflies per transform
    static_assert(T % C == 0 && NB % C == 0, "stage-B twiddle must be a per-thread constant");
    static_assert(SPLIT == 1 || (SPLIT == 2 && LPC == 1 && TP % 32 == 0), "the two halves of a line must be whole warps");
    static_assert(!PERSIST || LPC == 1, "the persistent variant handles one line per CTA at a time");
    CHB_DYN_SMEM(cplx, smem);
    const int tl = threadIdx.x % T, l = threadIdx.x / T;
    const int nx = g.nx, nxB = g.nxB, nzB = g.nzB;
    constexpr bool multi = MULTI;
    cplx* S = smem + (size_t)l * 6 * LB;
    // persistent variant: staging area of the next line's inputs [3][nx+1] and the twiddle tables, behind the six buffers
    const int nxp = nx + 1;
    cplx* const Xs = smem + 6 * LB;            // [3][nxp]
    cplx* const WhS = Xs + 3 * nxp;            // Wh[0..nx]
    cplx* const W1S = WhS + nxp;               // W[0..BC)
    cplx* const WbS = W1S + BC;                // W[A k], k = 0..B-1
    auto ldWh = [&](int i) -> cplx { if constexpr (PERSIST) return WhS[i]; else return __ldg(&Wh[i]); };
    auto ldW1 = [&](int t1) -> cplx { if constexpr (PERSIST) return W1S[t1]; else return __ldg(&W[t1]); };
    auto ldWb = [&](int k) -> cplx { if constexpr (PERSIST) return WbS[k]; else return __ldg(&W[A * k]); };

    // element (ka, t = b*C + c) of a transform buffer; for C = 4 the column index is swizzled so that
    // stage A (lanes = consecutive t), stage B (lanes = (ka, c)) and stage C (lanes = consecutive kb)
    // all hit 8 distinct 16-byte banks per quarter warp:  bank = (ka + b + 2c) mod 8
    auto sig = [](int t) -> int {
        if constexpr (C == 4) {
            const int bb = t >> 2, c = t & 3;
            return 8 * (c * (B / 8) + (bb >> 3)) + ((bb + 2 * c) & 7);
        } else {
            return t;
        }
    };
    // velocity buffer (transpose_index.h), 32-bit offsets: one GPU: row-major [comp][plane][z row][x];
    // several: [src rank][comp][plane][x tile][z row][x in tile] (row-major if g.twa < 0)
    const unsigned planeA = (unsigned)((size_t)nzB * nxB), cstrA = (unsigned)np * planeA;
    const bool tiledA = multi && g.twa >= 0;
    const unsigned tmaskA = tiledA ? (1u << g.twa) - 1u : 0u;
    const unsigned tstrA = tiledA ? ((unsigned)nzB << g.twa) : 0u;   // elements between x tiles
    auto row_of = [&](int izl_, int pli_) -> unsigned {
        return (unsigned)pli_ * planeA + (tiledA ? ((unsigned)izl_ << g.twa) : (unsigned)izl_ * (unsigned)nxB);
    };
    auto mode_off = [&](int n) -> unsigned {   // offset of mode n inside a (line, component)
        unsigned o = (unsigned)n;
        if (multi) {
            const unsigned qr = (unsigned)n / (unsigned)nxB, nl = (unsigned)n - qr * (unsigned)nxB;
            o = qr * 3u * cstrA + (tiledA ? ((nl >> g.twa) * tstrA + (nl & tmaskA)) : nl);
        }
        return o;
    };
    const int nlines = (nzB / LPC) * nplanes;
    auto prefetch = [&](int line) {   // inputs of `line` -> Xs (asynchronous)
        const unsigned rA = row_of(line % nzB, line / nzB);
        for (int e = threadIdx.x; e < 3 * nxp; e += T) {
            const int comp = e / nxp, n = e - comp * nxp;
            cp_async16(Xs + e, Ar + rA + (unsigned)comp * cstrA + mode_off(n));
        }
    };
    if constexpr (PERSIST) {
        for (int i = threadIdx.x; i < nxp; i += T) WhS[i] = Wh[i];
        for (int i = threadIdx.x; i < BC; i += T) W1S[i] = W[i];
        for (int i = threadIdx.x; i < B; i += T) WbS[i] = W[A * i];
        if ((int)blockIdx.x < nlines) prefetch(blockIdx.x);
    }
    auto do_line = [&](const int line) {
    const int izl = PERSIST ? line % nzB : (int)blockIdx.x * LPC + l;
    const int pli = PERSIST ? line / nzB : (int)blockIdx.y;
    const int iy = plane0 + pli - 1;
    const unsigned rowA = row_of(izl, pli);
    if constexpr (PERSIST) {
        cp_async_wait_all();
        __syncthreads();   // this line's inputs (and, the first time, the tables) are in shared memory
    }
    // ---- backward stage A: split pass -> radix-A -> smem; tasks = (component, mode group t1) --------
#pragma unroll 1
    for (int task = tl; task < 3 * BC; task += T) {
        const int comp = task / BC, t1 = task - comp * BC;
        const cplx wh1 = ldWh(t1);   // exp(+i pi t1 / M)
        const cplx w1 = ldW1(t1);    // exp(+2 pi i t1 / M)
        const cplx* __restrict__ Xc = Ar + rowA + (unsigned)comp * cstrA;
        auto X = [&](int n) -> cplx {   // mode n of this line, zero beyond nx (x zero-padding, dnsdata.f90:535)
            if (n > nx) return make_double2(0.0, 0.0);
            if constexpr (PERSIST) return Xs[comp * nxp + n];
            else return __ldg(Xc + mode_off(n));
        };
        cplx x[A];
        static_for<A>([&](auto a_) {
            constexpr int a = decltype(a_)::value;
            const int n = a * BC + t1;
            const cplx xa = X(n), xb = X(M - n);
            if (a == 0 && t1 == 0) {
                x[a] = make_double2(xa.x, xa.x);   // Z[0] = X0 + XM + i (X0 - XM), XM = 0, Im X0 ignored
            } else {
                const cplx s = make_double2(xa.x + xb.x, xa.y - xb.y);
                const cplx d = make_double2(xa.x - xb.x, xa.y + xb.y);
                const cplx t = cmul(mulw<2 * A, a, +1>(wh1), d);   // e^{i pi n/M} = e^{i pi t1/M} e^{i pi a/A}
                x[a] = make_double2(s.x - t.y, s.y + t.x);
            }
        });
        Dft<A, +1>::run(x);
        if (t1 != 0) apply_twiddle_seq<A>(x, w1);
        cplx* dst = S + comp * LB + sig(t1);
        static_for<A>([&](auto ka_) { constexpr int ka = decltype(ka_)::value; dst[ka * BCP] = x[ka]; });
    }
    __syncthreads();
    if constexpr (PERSIST) {   // the staging area is free: fetch the next line's inputs under the remaining stages
        if (line + (int)gridDim.x < nlines) prefetch(line + (int)gridDim.x);
    }
    // ---- backward stage B, in place; tasks = (component, ka, c) ---------------------------------------
    {
        const int cc = tl % C;
        const cplx w1 = ldWb(cc);   // w_BC^cc
#pragma unroll 1
        for (int task = tl; task < 3 * NB; task += T) {
            const int comp = task / NB, u = task - comp * NB;
            cplx* base = S + comp * LB + (u / C) * BCP;
            cplx x[B];
            static_for<B>([&](auto b_) { constexpr int b = decltype(b_)::value; x[b] = base[sig(b * C + cc)]; });
            Dft<B, +1>::run(x);
            if (cc != 0) apply_twiddle_seq<B>(x, w1);
            static_for<B>([&](auto b_) { constexpr int b = decltype(b_)::value; base[sig(b * C + cc)] = x[b]; });
        }
    }
    __syncthreads();
    // ---- backward stage C -> physical space -> CFL, products -> forward stage C --------------
    if constexpr (SPLIT == 1) {
        {
            // C = 4: consecutive lanes = consecutive kb (see sig); else consecutive ka
            const int ka = (C == 4) ? tl / B : tl % A, kb = (C == 4) ? tl % B : tl / A;
            cplx* base = S + ka * BCP;
            int sc_[C];
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; sc_[c] = sig(kb * C + c); });
            cplx U[C], V[C], Wv[C];
            static_for<C>([&](auto c_) {
                constexpr int c = decltype(c_)::value;
                U[c] = base[sc_[c]];
                V[c] = base[LB + sc_[c]];
                Wv[c] = base[2 * LB + sc_[c]];
            });
            Dft<C, +1>::run(U);
            Dft<C, +1>::run(V);
            Dft<C, +1>::run(Wv);
            if (compute_cfl) {   // dnsdata.f90:552-556 (block-uniform branch)
                double cmax = 0.0;
                if (iy >= 1 && iy <= g.ny - 1) {
                    const double rdx = 1.0 / g.dx, rdz = 1.0 / g.dz, rdy = 1.0 / dy[iy + 1];
                    static_for<C>([&](auto c_) {
                        constexpr int c = decltype(c_)::value;
                        cmax = fmax(cmax, fabs(U[c].x) * rdx + fabs(V[c].x) * rdy + fabs(Wv[c].x) * rdz);
                        cmax = fmax(cmax, fabs(U[c].y) * rdx + fabs(V[c].y) * rdy + fabs(Wv[c].y) * rdz);
                    });
                }
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
                if ((threadIdx.x & 31) == 0 && cmax > 0.0)
                    atomicMax(&sc->cfl_bits, (unsigned long long)__double_as_longlong(cmax));
            }
            const double f = 0.5 * g.factor;       // the 1/2 of the merge pass is folded into the products
            cplx wk1 = ldWb(kb);   // conj w_BC^kb
            wk1.y = -wk1.y;
            auto forward_c = [&](cplx* x, int p) {
                Dft<C, -1>::run(x);
                if (kb != 0) apply_twiddle_seq<C>(x, wk1);
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; base[p * LB + sc_[c]] = x[c]; });
            };
            cplx x[C];
            // slots (0..5) = (uu, vv, ww, uv, vw, uw) * factor                          dnsdata.f90:581-584
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(U[c].x * V[c].x * f, U[c].y * V[c].y * f); });
            forward_c(x, 3);
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(V[c].x * Wv[c].x * f, V[c].y * Wv[c].y * f); });
            forward_c(x, 4);
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(U[c].x * Wv[c].x * f, U[c].y * Wv[c].y * f); });
            forward_c(x, 5);
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; U[c] = make_double2(U[c].x * U[c].x * f, U[c].y * U[c].y * f); });
            forward_c(U, 0);
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; V[c] = make_double2(V[c].x * V[c].x * f, V[c].y * V[c].y * f); });
            forward_c(V, 1);
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; Wv[c] = make_double2(Wv[c].x * Wv[c].x * f, Wv[c].y * Wv[c].y * f); });
            forward_c(Wv, 2);
        }
    } else {
        {
            const int pos = tl % TP, half = tl / TP;      // half is warp-uniform (TP is a multiple of 32)
            // C = 4: consecutive lanes = consecutive kb (see sig); else consecutive ka
            const int ka = (C == 4) ? pos / B : pos % A, kb = (C == 4) ? pos % B : pos / A;
            cplx* base = S + ka * BCP;
            int sc_[C];
            static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; sc_[c] = sig(kb * C + c); });
            cplx U[C], V[C], Wv[C];
            static_for<C>([&](auto c_) {
                constexpr int c = decltype(c_)::value;
                U[c] = base[sc_[c]];
                V[c] = base[LB + sc_[c]];
                Wv[c] = base[2 * LB + sc_[c]];
            });
            __syncthreads();   // both threads of a position have read u, v, w before either overwrites them
            Dft<C, +1>::run(U);
            Dft<C, +1>::run(V);
            Dft<C, +1>::run(Wv);
            if (compute_cfl && half == 0) {   // dnsdata.f90:552-556 (warp-uniform branch)
                double cmax = 0.0;
                if (iy >= 1 && iy <= g.ny - 1) {
                    const double rdx = 1.0 / g.dx, rdz = 1.0 / g.dz, rdy = 1.0 / dy[iy + 1];
                    static_for<C>([&](auto c_) {
                        constexpr int c = decltype(c_)::value;
                        cmax = fmax(cmax, fabs(U[c].x) * rdx + fabs(V[c].x) * rdy + fabs(Wv[c].x) * rdz);
                        cmax = fmax(cmax, fabs(U[c].y) * rdx + fabs(V[c].y) * rdy + fabs(Wv[c].y) * rdz);
                    });
                }
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
                if ((threadIdx.x & 31) == 0 && cmax > 0.0)
                    atomicMax(&sc->cfl_bits, (unsigned long long)__double_as_longlong(cmax));
            }
            const double f = 0.5 * g.factor;       // the 1/2 of the merge pass is folded into the products
            cplx wk1 = ldWb(kb);   // conj w_BC^kb
            wk1.y = -wk1.y;
            auto forward_c = [&](cplx* x, int p) {
                Dft<C, -1>::run(x);
                if (kb != 0) apply_twiddle_seq<C>(x, wk1);
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; base[p * LB + sc_[c]] = x[c]; });
            };
            cplx x[C];
            // slots (0..5) = (uu, vv, ww, uv, vw, uw) * factor                          dnsdata.f90:581-584
            if (half == 0) {
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(U[c].x * U[c].x * f, U[c].y * U[c].y * f); });
                forward_c(x, 0);
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(V[c].x * V[c].x * f, V[c].y * V[c].y * f); });
                forward_c(x, 1);
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(Wv[c].x * Wv[c].x * f, Wv[c].y * Wv[c].y * f); });
                forward_c(x, 2);
            } else {
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(U[c].x * V[c].x * f, U[c].y * V[c].y * f); });
                forward_c(x, 3);
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(V[c].x * Wv[c].x * f, V[c].y * Wv[c].y * f); });
                forward_c(x, 4);
                static_for<C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = make_double2(U[c].x * Wv[c].x * f, U[c].y * Wv[c].y * f); });
                forward_c(x, 5);
            }
        }
    }
    __syncthreads();
    // ---- forward stage B, in place; tasks = (product, ka, c) ------------------------------------------
#pragma unroll 1
    for (int task = tl; task < 6 * NB; task += T) {
        const int p = task / NB, u = task - p * NB;
        cplx* base = S + p * LB + (u / C) * BCP;
        const int cc = u % C;
        cplx x[B];
        static_for<B>([&](auto b_) { constexpr int b = decltype(b_)::value; x[b] = base[sig(b * C + cc)]; });
        Dft<B, -1>::run(x);
        static_for<B>([&](auto b_) { constexpr int b = decltype(b_)::value; base[sig(b * C + cc)] = x[b]; });
    }
    __syncthreads();
    // ---- forward stage A: twiddle, radix-A, natural order back to smem; tasks = (product, t1) ------
#pragma unroll 1
    for (int task = tl; task < 6 * BC; task += T) {
        const int p = task / BC, t1 = task - p * BC;
        cplx w1 = ldW1(t1);
        w1.y = -w1.y;
        cplx* col = S + p * LB + sig(t1);
        cplx x[A];
        static_for<A>([&](auto ka_) { constexpr int ka = decltype(ka_)::value; x[ka] = col[ka * BCP]; });
        if (t1 != 0) apply_twiddle_seq<A>(x, w1);
        Dft<A, -1>::run(x);
        static_for<A>([&](auto a_) { constexpr int a = decltype(a_)::value; col[a * BCP] = x[a]; });   // Z[a*BC + t1]
    }
    __syncthreads();
    // ---- merge pass + x-dealiasing (keep modes 0..nx) + store ----------------------------------
    for (int j = tl; j <= nx; j += T) {
        const int pj = (j / BC) * BCP + sig(j % BC);
        const int jm = (j == 0) ? 0 : M - j;
        const int pm = (jm / BC) * BCP + sig(jm % BC);
        cplx w = ldWh(j);
        w.y = -w.y;   // e^{-i pi j/M}
        const int q = multi ? j / nxB : 0;
        cplx* __restrict__ Bout = multi ? Bw.p[q] : Bw.p[0];   // the owner of x-mode j (this GPU's or a peer's HBM over NVLink)
        const size_t o0 = chb_bufB_index(g.rank, 6, 0, np, pli, nzB, izl, nxB, j - q * nxB, g.tw);
        const size_t ostride = (size_t)np * nzB * nxB;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const cplx z = S[p * LB + pj];
            const cplx zm = S[p * LB + pm];
            const cplx e = make_double2(z.x + zm.x, z.y - zm.y);   // (Z + conj Zm)/2, the 1/2 is in the products
            const cplx d = make_double2(z.x - zm.x, z.y + zm.y);   // (Z - conj Zm)/2
            const cplx o = make_double2(d.y, -d.x);                                // -i * d
            Bout[o0 + p * ostride] = cadd(e, cmul(w, o));
        }
    }
    if constexpr (PERSIST) __syncthreads();   // the buffers are rewritten by the next line's stage A
    };   // do_line
    if constexpr (PERSIST) {
#pragma unroll 1
        for (int line = (int)blockIdx.x; line < nlines; line += (int)gridDim.x) do_line(line);
    } else {
        do_line(0);
    }
}

#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)   // the kernel-only emulation harnesses (tests/host_emul) stop here
// SMs the stream the conv launchers use can run on (its green-context partition, else the device)
static int stream_sms(chb_handle_s* h) {
    if (h->green_sms[0] && h->cstream == h->sA) return h->green_sms[0];
    if (h->green_sms[1] && h->cstream == h->sB) return h->green_sms[1];
    static int dev_sms = 0;
    if (!dev_sms) cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device);
    return dev_sms > 0 ? dev_sms : 148;
}

// SPLIT = 1 | 2 threads per innermost butterfly position; PERSIST: persistent CTAs with prefetched inputs (LPC = 1)
template <class G, int LPC, int MINB, int SPLIT, bool PERSIST>
static bool launch_x4(chb_handle_s* h, int plane0, int nplanes, int compute_cfl) {
    constexpr int T = SPLIT * (G::N / G::C);
    constexpr int LB = G::A * (G::BC + 1);
    if (h->g.nzB % LPC != 0) return false;
    size_t smem = (size_t)LPC * 6 * LB * sizeof(cplx);
    if (PERSIST) smem += ((size_t)4 * (h->g.nx + 1) + G::BC + G::B) * sizeof(cplx);
    if (smem > 227 * 1024) return false;
    auto kern = (h->g.nranks > 1) ? xpass4_kernel<G, LPC, MINB, true, SPLIT, PERSIST> : xpass4_kernel<G, LPC, MINB, false, SPLIT, PERSIST>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(h->g.nzB / LPC, nplanes);
    if (PERSIST) {
        const long long nlines = (long long)h->g.nzB * nplanes, ncta = (long long)stream_sms(h) * MINB;
        grid = dim3((unsigned)(nlines < ncta ? nlines : ncta), 1);
    }
    ScopedKernelTimer tm(h, "xpass", h->cstream);
    CHB_LAUNCH(grid, LPC * T, smem, h->cstream, kern)(h->Ar, h->Bw, h->g, h->Wx, h->Wh, h->t_dy, h->sc, plane0, h->chunk_planes,
                                             compute_cfl, nplanes);
    h->launches++;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Variants measured on B200 (profiles/r2a_variants.md, r2e_xpass_persist.md): nxd = 768: one thread per position, three
// CTAs per SM; nxd = 1536: two threads per position (CHB_XPASS_SPLIT), persistent with prefetch (CHB_XPASS_PERSIST).
bool launch_x3_pass(chb_handle_s* h, int plane0, int nplanes, int compute_cfl) {
    const int sp = h->xpass_split, pe = h->xpass_persist;
    switch (h->g.nxd) {
        case 384: return launch_x4<Fft3<384, 12, 8, 4>, 1, 6, 1, false>(h, plane0, nplanes, compute_cfl);
        case 768:
            if (pe && sp) return launch_x4<Fft3<768, 12, 16, 4>, 1, 2, 2, true>(h, plane0, nplanes, compute_cfl);
            if (pe) return launch_x4<Fft3<768, 12, 16, 4>, 1, 2, 1, true>(h, plane0, nplanes, compute_cfl);
            if (sp) return launch_x4<Fft3<768, 12, 16, 4>, 1, 2, 2, false>(h, plane0, nplanes, compute_cfl);
            return launch_x4<Fft3<768, 12, 16, 4>, 1, 3, 1, false>(h, plane0, nplanes, compute_cfl);
        case 1536:
            if (pe && sp) return launch_x4<Fft3<1536, 12, 16, 8>, 1, 1, 2, true>(h, plane0, nplanes, compute_cfl);
            if (pe) return launch_x4<Fft3<1536, 12, 16, 8>, 1, 1, 1, true>(h, plane0, nplanes, compute_cfl);
            if (sp) return launch_x4<Fft3<1536, 12, 16, 8>, 1, 1, 2, false>(h, plane0, nplanes, compute_cfl);
            return launch_x4<Fft3<1536, 12, 16, 8>, 1, 1, 1, false>(h, plane0, nplanes, compute_cfl);
        default: return false;
    }
}
#endif
