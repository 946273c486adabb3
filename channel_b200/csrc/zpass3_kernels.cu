// zpass3_kernels.cu - z-direction passes of the nonlinear term for the large transform sizes,
// built on the register-resident three-stage FFT of fft_regs.cuh.
//
//   zfwd3: zero-pad in z + backward complex FFT of length nzd      (dnsdata.f90:504-510, IFT ffts.f90:71)
//          + zTOx pack: output x-contiguous, one block per destination rank (mpi_transpose.f90:64-71)
//   zbwd3: xTOz unpack + forward complex FFT of length nzd (FFT ffts.f90:70) + z-truncation
//          through izd() (DD macro, dnsdata.f90:609)
//
// One CTA = LPC neighbouring x-modes (lines), TPL threads per line.  The two stages that touch
// the z-contiguous side run "line-major"; the stage that touches the work buffer runs
// "cross-line" (consecutive lanes = consecutive lines, then consecutive z rows), and the work
// buffers are tiled [x tile of LPC][z row][x in tile] (transpose_index.h) so that a warp's
// accesses are contiguous.  Per point: one global read, one global write, two shared-memory
// exchanges.
#include "chb_internal.h"
#include "fft_regs.cuh"

// x[p] *= w1^p, p = 1..R-1, sequential recurrence (low register pressure)
template <int R>
__device__ __forceinline__ void z_twiddle_seq(cplx* x, cplx w1) {
    cplx wp = w1;
    x[1] = cmul(x[1], wp);
    static_for<R - 2>([&](auto i_) {
        constexpr int p = decltype(i_)::value + 2;
        wp = cmul(wp, w1);
        x[p] = cmul(x[p], wp);
    });
}

// ---------------------------------------------------------------------------------------------
// zfwd4 / zbwd4: the same three stages with TPL (a multiple of 32) threads per line and
// block-wide barriers between stages, sized so that MINB CTAs are resident per SM: while one CTA
// waits for its lines to arrive from HBM another one computes.
// DIRECT (CHB_ZF_DIRECT=1; measured slower, 25.7 against 21.3 ms/step at config 3, kept as a comparator): stage A reads its inputs
// straight from global memory (a warp reads 512 contiguous bytes per radix digit) instead of through the TMA
// staging copy: two of the six shared-memory passes per point disappear, the load latency moves into the threads.
template <class G, int LPC, int TPL, int MINB, bool DIRECT = false>
__global__ void __launch_bounds__(LPC * TPL, MINB)
zfwd4_kernel(const cplx* __restrict__ V, PeerPtrs Aw, Geometry g, const cplx* __restrict__ W, int plane0, int np,
             int LS) {
    CHB_DYN_SMEM(cplx, smem);
    constexpr int BCP = G::BC + 1;
    static_assert(TPL % G::C == 0, "stage-B twiddle must be a per-thread constant");
    const int tl = threadIdx.x % TPL, wl = threadIdx.x / TPL;
    const int ixl0 = blockIdx.x * LPC;
    const int pli = blockIdx.y, comp = blockIdx.z;
    const int iyp = plane0 + pli;
    const int nz = g.nz;
    __shared__ unsigned long long mbar[LPC];
    // the twiddles this thread needs first, loaded while the line is on its way from HBM (they are L1 / L2 hits, but
    // their latency sat right behind the barriers: 17-36 % of the stall samples of stages A and B).  An L2 prefetch
    // (cp.async.bulk.prefetch.L2) of the lines of the CTA that runs 300-2400 CTAs later was measured too and rejected:
    // zbwd 45.6 -> 47.8-50.6 ms/step at config 3 (profiles/r2c_r2e_single_gpu.md)
    cplx wa_next = ctw<+1>(W, tl);
    const cplx wb1 = ctw<+1>(W, G::A * (tl % G::C));
    {   // ---- stage A, line-major; the V line is staged by TMA bulk copies into the in-place layout
        const cplx* __restrict__ src = V + (((size_t)comp * g.nyp + iyp) * g.nxB + ixl0 + wl) * g.nzt;
        cplx* sm = smem + wl * LS;
        if constexpr (!DIRECT) {
            if (tl == 0) {
                mbar_init(&mbar[wl], 1);
                mbar_expect_tx(&mbar[wl], (unsigned)(g.nzt * sizeof(cplx)));
                for (int a = 0; a < G::A; ++a) {
                    const int lo = a * G::BC, hi = lo + G::BC - 1;
                    const int h1 = min(hi, nz);                    // rows 1..nz+1        <- V(iy,0:nz)
                    if (h1 >= lo) bulk_g2s(sm + a * BCP, src + nz + lo, (unsigned)((h1 - lo + 1) * sizeof(cplx)), &mbar[wl]);
                    const int l2 = max(lo, G::N - nz);             // rows nzd-nz+1..nzd  <- V(iy,-nz:-1)
                    if (hi >= l2)
                        bulk_g2s(sm + a * BCP + (l2 - lo), src + (l2 - (G::N - nz)), (unsigned)((hi - l2 + 1) * sizeof(cplx)),
                                 &mbar[wl]);
                }
            }
            __syncthreads();
            mbar_wait(&mbar[wl], 0);
        }
#pragma unroll 1
        for (int t1 = tl; t1 < G::BC; t1 += TPL) {
            const cplx wa = wa_next;
            if (t1 + TPL < G::BC) wa_next = ctw<+1>(W, t1 + TPL);   // next iteration's twiddle under this one's butterfly
            cplx x[G::A];
            static_for<G::A>([&](auto a_) {
                constexpr int a = decltype(a_)::value;
                const int n = a * G::BC + t1;
                if constexpr (DIRECT)
                    x[a] = (n <= nz) ? __ldg(src + nz + n) : ((n >= G::N - nz) ? __ldg(src + (n - (G::N - nz))) : make_double2(0.0, 0.0));
                else
                    x[a] = (n <= nz || n >= G::N - nz) ? sm[a * BCP + t1] : make_double2(0.0, 0.0);
            });
            Dft<G::A, +1>::run(x);
            if (t1 != 0) z_twiddle_seq<G::A>(x, wa);
            static_for<G::A>([&](auto ka_) {
                constexpr int ka = decltype(ka_)::value;
                sm[ka * BCP + t1] = x[ka];
            });
        }
    }
    __syncthreads();
    {   // ---- stage B, line-major, in place
        cplx* sm = smem + wl * LS;
        const int cc = tl % G::C;
        const cplx w1 = wb1;
#pragma unroll 1
        for (int u = tl; u < G::A * G::C; u += TPL) {
            cplx* base = sm + (u / G::C) * BCP + cc;
            cplx x[G::B];
            static_for<G::B>([&](auto b_) { constexpr int b = decltype(b_)::value; x[b] = base[b * G::C]; });
            Dft<G::B, +1>::run(x);
            if (cc != 0) z_twiddle_seq<G::B>(x, w1);
            static_for<G::B>([&](auto b_) { constexpr int b = decltype(b_)::value; base[b * G::C] = x[b]; });
        }
    }
    __syncthreads();
    {   // ---- stage C, cross-line: smem -> registers -> global (x-contiguous, per destination rank)
        const int l = threadIdx.x % LPC, q = threadIdx.x / LPC;
        const cplx* sm = smem + l * LS;
        const int nzB = g.nzB, nxB = g.nxB;
        const bool multi = g.nranks > 1;
#pragma unroll 1
        for (int t = q; t < G::AB; t += TPL) {
            const cplx* base = sm + (t % G::A) * BCP + (t / G::A) * G::C;
            cplx x[G::C];
            static_for<G::C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = base[c]; });
            Dft<G::C, +1>::run(x);
            static_for<G::C>([&](auto kc_) {
                constexpr int kc = decltype(kc_)::value;
                const int k = t + G::AB * kc;
                const int peer = multi ? k / nzB : 0;
                cplx* dst = multi ? Aw.p[peer] : Aw.p[0];
                dst[chb_bufA_index(g.rank, comp, np, pli, nzB, k - peer * nzB, nxB, ixl0 + l, g.twa)] = x[kc];
            });
        }
    }
}

template <class G, int LPC, int TPL, int MINB>
__global__ void __launch_bounds__(LPC * TPL, MINB)
zbwd4_kernel(const cplx* __restrict__ Br, cplx* __restrict__ P, Geometry g, const cplx* __restrict__ W, int plane0, int np,
             int LS) {
    CHB_DYN_SMEM(cplx, smem);
    constexpr int BCP = G::BC + 1;
    static_assert(TPL % G::C == 0, "stage-B twiddle must be a per-thread constant");
    const int tl = threadIdx.x % TPL, wl = threadIdx.x / TPL;
    const int ixl0 = blockIdx.x * LPC;
    const int pli = blockIdx.y, comp = blockIdx.z;  // product index 0..5
    const int nz = g.nz;
    (void)plane0;
    cplx wa_next = ctw<-1>(W, tl);                              // see zfwd4
    const cplx wb1 = ctw<-1>(W, G::A * (tl % G::C));
    {   // ---- staging, cross-line: per-thread 16-byte cp.async straight into the in-place layout
        const int l = threadIdx.x % LPC, q = threadIdx.x / LPC;
        cplx* sml = smem + l * LS;
        const int nzB = g.nzB, nxB = g.nxB;
        const bool multi = g.nranks > 1;
#pragma unroll 4
        for (int n = q; n < G::N; n += TPL) {
            const int peer = multi ? n / nzB : 0;
            cp_async16(sml + (n / G::BC) * BCP + (n % G::BC),
                       Br + chb_bufB_index(peer, 6, comp, np, pli, nzB, n - peer * nzB, nxB, ixl0 + l, g.tw));
        }
        cp_async_wait_all();
    }
    __syncthreads();
    cplx* sm = smem + wl * LS;
    // ---- stage A, line-major, in place
#pragma unroll 1
    for (int t1 = tl; t1 < G::BC; t1 += TPL) {
        const cplx wa = wa_next;
        if (t1 + TPL < G::BC) wa_next = ctw<-1>(W, t1 + TPL);
        cplx x[G::A];
        static_for<G::A>([&](auto a_) { constexpr int a = decltype(a_)::value; x[a] = sm[a * BCP + t1]; });
        Dft<G::A, -1>::run(x);
        if (t1 != 0) z_twiddle_seq<G::A>(x, wa);
        static_for<G::A>([&](auto ka_) { constexpr int ka = decltype(ka_)::value; sm[ka * BCP + t1] = x[ka]; });
    }
    __syncthreads();
    {   // ---- stage B, line-major, in place
        const int cc = tl % G::C;
        const cplx w1 = wb1;
#pragma unroll 1
        for (int u = tl; u < G::A * G::C; u += TPL) {
            cplx* base = sm + (u / G::C) * BCP + cc;
            cplx x[G::B];
            static_for<G::B>([&](auto b_) { constexpr int b = decltype(b_)::value; x[b] = base[b * G::C]; });
            Dft<G::B, -1>::run(x);
            if (cc != 0) z_twiddle_seq<G::B>(x, w1);
            static_for<G::B>([&](auto b_) { constexpr int b = decltype(b_)::value; base[b * G::C] = x[b]; });
        }
    }
    __syncthreads();
    {   // ---- stage C, line-major: smem -> registers -> global (z-contiguous, truncated to -nz..nz)
        // P = the chunk's spectral products [6][np][nxB][2nz+1]
        cplx* __restrict__ dst = P + (((size_t)comp * np + pli) * g.nxB + ixl0 + wl) * g.nzt;
#pragma unroll 1
        for (int t = tl; t < G::AB; t += TPL) {
            const cplx* base = sm + (t % G::A) * BCP + (t / G::A) * G::C;
            cplx x[G::C];
            static_for<G::C>([&](auto c_) { constexpr int c = decltype(c_)::value; x[c] = base[c]; });
            Dft<G::C, -1>::run(x);
            static_for<G::C>([&](auto kc_) {
                constexpr int kc = decltype(kc_)::value;
                const int k = t + G::AB * kc;                     // izd(iz) = k  (dnsdata.f90:156)
                if (k <= nz) dst[nz + k] = x[kc];
                else if (k >= G::N - nz) dst[k - (G::N - nz)] = x[kc];
            });
        }
    }
}

#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)   // the kernel-only emulation harnesses (tests/host_emul) stop here
template <class G, int LPC, int TPL, int MINB>
static bool launch_z4(chb_handle_s* h, int plane0, int nplanes, bool fwd) {
    constexpr int BCP = G::BC + 1;
    int LS = G::A * BCP;
    const int want = (LPC == 8) ? 1 : (LPC == 4 ? 2 : 4);   // line stride mod 8 that keeps the cross-line accesses conflict-free
    while (LS % 8 != want) ++LS;
    const size_t smem = (size_t)LPC * LS * sizeof(cplx);
    if (h->g.nxB % LPC != 0) return false;
    dim3 grid(h->g.nxB / LPC, nplanes, fwd ? 3 : 6);
    if (fwd) {
        auto kern = h->zf_direct ? zfwd4_kernel<G, LPC, TPL, MINB, true> : zfwd4_kernel<G, LPC, TPL, MINB, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        ScopedKernelTimer tm(h, "zfwd", h->cstream);
        CHB_LAUNCH(grid, LPC * TPL, smem, h->cstream, kern)(h->V, h->Aw, h->g, h->Wz, plane0, h->chunk_planes, LS);
    } else {
        cudaFuncSetAttribute(zbwd4_kernel<G, LPC, TPL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(zbwd4_kernel<G, LPC, TPL, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        ScopedKernelTimer tm(h, "zbwd", h->cstream);
        CHB_LAUNCH(grid, LPC * TPL, smem, h->cstream, zbwd4_kernel<G, LPC, TPL, MINB>)(h->Br, h->Pc, h->g, h->Wz, plane0, h->chunk_planes, LS);
    }
    h->launches++;
    return true;
}

// ---------------------------------------------------------------------------------------------
// returns false when no specialised kernel exists for this size (caller falls back to the
// generic shared-memory passes of conv_kernels.cu)
bool launch_z3_fwd_or_bwd(chb_handle_s* h, int plane0, int nplanes, bool fwd) {
    if (h->g.nz < 2) return false;
    const int lpc = fwd ? h->zf_lines_per_cta : h->zb_lines_per_cta;
    // CHB_Z_TPL=128 | 96 (measured, profiles/r2a_variants.md): more threads per line, i.e. more
    // resident warps per SM at the same shared memory, with two lines per CTA:
    //   nzd = 3072: 128 threads/line, 2 CTAs/SM -> 16 warps (default 8), 128 registers, 0 / 8 bytes of spills
    //   nzd = 1536: 128 threads/line, 3 CTAs/SM -> 24 warps (default 16), 80 registers, no spills
    //               96 threads/line, 4 CTAs/SM -> 24 warps, 80 registers, no spills
    // measured (profiles/r2a_variants.md): 128 threads per line win for the backward pass at nzd = 3072 (7.2 against 7.5 ms),
    // lose everywhere else; CHB_Z_TPL=64 forces the 64-thread kernels
    if ((h->z_tpl == 128 || (h->z_tpl == 0 && !fwd)) && lpc == 2 && h->g.nzd == 3072)
        return launch_z4<Fft3<3072, 12, 16, 16>, 2, 128, 2>(h, plane0, nplanes, fwd);
    if (h->z_tpl == 128 && lpc == 2 && h->g.nzd == 1536) return launch_z4<Fft3<1536, 12, 16, 8>, 2, 128, 3>(h, plane0, nplanes, fwd);
    if (h->z_tpl == 96 && lpc == 2 && h->g.nzd == 1536) return launch_z4<Fft3<1536, 12, 16, 8>, 2, 96, 4>(h, plane0, nplanes, fwd);
    switch (h->g.nzd * 16 + lpc) {
        case 768 * 16 + 2: return launch_z4<Fft3<768, 12, 8, 8>, 2, 64, 8>(h, plane0, nplanes, fwd);
        case 768 * 16 + 4: return launch_z4<Fft3<768, 12, 8, 8>, 4, 64, 4>(h, plane0, nplanes, fwd);
        case 768 * 16 + 8: return launch_z4<Fft3<768, 12, 8, 8>, 8, 32, 2>(h, plane0, nplanes, fwd);
        case 1536 * 16 + 2: return launch_z4<Fft3<1536, 12, 16, 8>, 2, 64, 4>(h, plane0, nplanes, fwd);
        case 1536 * 16 + 4: return launch_z4<Fft3<1536, 12, 16, 8>, 4, 64, 2>(h, plane0, nplanes, fwd);
        case 1536 * 16 + 8: return launch_z4<Fft3<1536, 12, 16, 8>, 8, 32, 1>(h, plane0, nplanes, fwd);
        case 3072 * 16 + 2: return launch_z4<Fft3<3072, 12, 16, 16>, 2, 64, 2>(h, plane0, nplanes, fwd);
        case 3072 * 16 + 4: return launch_z4<Fft3<3072, 12, 16, 16>, 4, 64, 1>(h, plane0, nplanes, fwd);
        default: return false;
    }
}
#endif
