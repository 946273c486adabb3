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
// SPLIT = 2 (experimental, CHB_XPASS_SPLIT=1; proven on the CPU emulator, not yet measured on a GPU): TWO threads per
// innermost butterfly position (T = 2 M/C threads per line).  For M = 1536, whose six line buffers (148.6 KB) leave
// room for one CTA per SM only, that is twice the resident warps at the same shared memory (12 instead of 6, 140
// registers); for M = 768, two CTAs of 12 warps at 80 registers instead of three of 6 warps at 96.  The flattened
// task loops of the outer stages take the extra threads as they are; in the innermost stage both threads of a
// position run the radix-C butterflies of u, v, w (duplicated: 3 of the 9 innermost butterflies per position), then
// one forms uu, vv, ww and the other uv, vw, uw.  Because two threads now read the u, v, w entries that the products
// overwrite in place, a barrier separates the reads from the writes.
template <class G, int LPC, int MINB, bool MULTI, int SPLIT = 1>
__global__ void __launch_bounds__(SPLIT * LPC * (G::N / G::C), MINB)
xpass4_kernel(const cplx* __restrict__ Ar, const __grid_constant__ PeerPtrs Bw, const __grid_constant__ Geometry g, const cplx* __restrict__ W,
              const cplx* __restrict__ Wh, const double* __restrict__ dy, DevScalars* sc, int plane0, int np,
              int compute_cfl) {
    constexpr int M = G::N, A = G::A, B = G::B, C = G::C, BC = G::BC;
    constexpr int TP = M / C;          // butterfly positions of the innermost stage
    constexpr int T = SPLIT * TP;      // threads per line
    constexpr int BCP = BC + 1;
    constexpr int LB = A * BCP;        // complex per buffer
    constexpr int NB = A * C;          // stage-B butterflies per transform
    static_assert(T % C == 0 && NB % C == 0, "stage-B twiddle must be a per-thread constant");
    static_assert(SPLIT == 1 || (SPLIT == 2 && LPC == 1 && TP % 32 == 0), "the two halves of a line must be whole warps");
    CHB_DYN_SMEM(cplx, smem);
    const int tl = threadIdx.x % T, l = threadIdx.x / T;
    const int izl = blockIdx.x * LPC + l;
    const int pli = blockIdx.y;
    const int iy = plane0 + pli - 1;
    const int nx = g.nx, nxB = g.nxB, nzB = g.nzB;
    constexpr bool multi = MULTI;
    cplx* S = smem + (size_t)l * 6 * LB;

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
    const unsigned rowA = (unsigned)pli * planeA + (tiledA ? ((unsigned)izl << g.twa) : (unsigned)izl * (unsigned)nxB);
    const unsigned tmaskA = tiledA ? (1u << g.twa) - 1u : 0u;
    const unsigned tstrA = tiledA ? ((unsigned)nzB << g.twa) : 0u;   // elements between x tiles
    // ---- backward stage A: split pass -> radix-A -> smem; tasks = (component, mode group t1) --------
#pragma unroll 1
    for (int task = tl; task < 3 * BC; task += T) {
        const int comp = task / BC, t1 = task - comp * BC;
        const cplx wh1 = Wh[t1];   // exp(+i pi t1 / M)
        const cplx w1 = W[t1];     // exp(+2 pi i t1 / M)
        const cplx* __restrict__ Xc = Ar + rowA + (unsigned)comp * cstrA;
        auto X = [&](int n) -> cplx {   // mode n of this line, zero beyond nx (x zero-padding, dnsdata.f90:535)
            if (n > nx) return make_double2(0.0, 0.0);
            unsigned o = (unsigned)n;
            if (multi) {
                const unsigned qr = (unsigned)n / (unsigned)nxB, nl = (unsigned)n - qr * (unsigned)nxB;
                o = qr * 3u * cstrA + (tiledA ? ((nl >> g.twa) * tstrA + (nl & tmaskA)) : nl);
            }
            return __ldg(Xc + o);
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
    // ---- backward stage B, in place; tasks = (component, ka, c) ---------------------------------------
    {
        const int cc = tl % C;
        const cplx w1 = W[A * cc];   // w_BC^cc
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
            const cplx wk1 = ctw<-1>(W, A * kb);   // conj w_BC^kb
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
            const cplx wk1 = ctw<-1>(W, A * kb);   // conj w_BC^kb
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
        const cplx w1 = ctw<-1>(W, t1);
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
        cplx w = Wh[j];
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
}

#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)   // the kernel-only emulation harnesses (tests/host_emul) stop here
template <class G, int LPC, int MINB>
static bool launch_x4(chb_handle_s* h, int plane0, int nplanes, int compute_cfl) {
    constexpr int T = G::N / G::C;
    constexpr int LB = G::A * (G::BC + 1);
    if (h->g.nzB % LPC != 0) return false;
    const size_t smem = (size_t)LPC * 6 * LB * sizeof(cplx);
    auto kern = (h->g.nranks > 1) ? xpass4_kernel<G, LPC, MINB, true> : xpass4_kernel<G, LPC, MINB, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(h->g.nzB / LPC, nplanes);
    ScopedKernelTimer tm(h, "xpass", h->cstream);
    CHB_LAUNCH(grid, LPC * T, smem, h->cstream, kern)(h->Ar, h->Bw, h->g, h->Wx, h->Wh, h->t_dy, h->sc, plane0, h->chunk_planes,
                                             compute_cfl);
    h->launches++;
    return true;
}


template <class G, int MINB>
static bool launch_x5(chb_handle_s* h, int plane0, int nplanes, int compute_cfl) {
    constexpr int T = 2 * (G::N / G::C);
    constexpr int LB = G::A * (G::BC + 1);
    const size_t smem = (size_t)6 * LB * sizeof(cplx);
    auto kern = (h->g.nranks > 1) ? xpass4_kernel<G, 1, MINB, true, 2> : xpass4_kernel<G, 1, MINB, false, 2>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(h->g.nzB, nplanes);
    ScopedKernelTimer tm(h, "xpass", h->cstream);
    CHB_LAUNCH(grid, T, smem, h->cstream, kern)(h->Ar, h->Bw, h->g, h->Wx, h->Wh, h->t_dy, h->sc, plane0, h->chunk_planes, compute_cfl);
    h->launches++;
    return true;
}

// ---------------------------------------------------------------------------------------------
bool launch_x3_pass(chb_handle_s* h, int plane0, int nplanes, int compute_cfl) {
    if (h->xpass_split && h->g.nxd == 1536) return launch_x5<Fft3<1536, 12, 16, 8>, 1>(h, plane0, nplanes, compute_cfl);
    if (h->xpass_split && h->g.nxd == 768) return launch_x5<Fft3<768, 12, 16, 4>, 2>(h, plane0, nplanes, compute_cfl);
    switch (h->g.nxd) {
        case 384: return launch_x4<Fft3<384, 12, 8, 4>, 1, 6>(h, plane0, nplanes, compute_cfl);
        case 768: return launch_x4<Fft3<768, 12, 16, 4>, 1, 3>(h, plane0, nplanes, compute_cfl);
        case 1536: return launch_x4<Fft3<1536, 12, 16, 8>, 1, 1>(h, plane0, nplanes, compute_cfl);
        default: return false;
    }
}
#endif
