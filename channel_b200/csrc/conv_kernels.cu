// conv_kernels.cu - the pseudo-spectral nonlinear term (convolutions, dnsdata.f90:487-602)
// as three batched kernels over a chunk of y-planes:
//
//   zfwd  : zero-pad in z + backward complex FFT of length nzd   (dnsdata.f90:504-510, IFT ffts.f90:71)
//           output already transposed to x-lines and split per destination rank (zTOx pack,
//           mpi_transpose.f90:64-71)
//   xpass : zero-pad in x + c2r (RFT ffts.f90:72) -> CFL (dnsdata.f90:552-556) -> the six
//           products * factor (dnsdata.f90:581-584) -> r2c (HFT ffts.f90:74) -> keep modes 0..nx
//           (xTOz pack, mpi_transpose.f90:99-106).  Physical-space data never leaves shared memory.
//   zbwd  : forward complex FFT of length nzd (FFT ffts.f90:70) + z-truncation through izd()
//           (DD macro, dnsdata.f90:609)
//
// Work-buffer layout (A: 3 components, B: 6 products), one contiguous block per peer rank so
// the pencil transposes are plain block exchanges:
//   buf[peer][comp][plane][izl][ixl]   izl in [0,nzB), ixl in [0,nxB)
// on the z side peer = owner of the physical z-line (iz_d / nzB); on the x side peer = owner of
// the x-mode (ix / nxB).  With one rank this is buf[comp][plane][iz_d][ix].
#include "chb_internal.h"

#define CONV_THREADS 256

#define buf_index chb_buf_index

// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CONV_THREADS)
zfwd_kernel(const cplx* __restrict__ V, PeerPtrs Aw, Geometry g, FftPlan pl, const cplx* __restrict__ W,
            const int* __restrict__ rev, int plane0, int np, int tx, int line_stride) {
    CHB_DYN_SMEM(cplx, smem);
    const int ixl0 = blockIdx.x * tx;
    const int pli = blockIdx.y;
    const int c = blockIdx.z;
    const int iyp = plane0 + pli;
    const int nl = min(tx, g.nxB - ixl0);
    const int nzd = g.nzd, nz = g.nz, nzt = g.nzt;
    const cplx* src = V + (((size_t)c * g.nyp + iyp) * g.nxB + ixl0) * nzt;
    for (int idx = threadIdx.x; idx < nl * nzd; idx += blockDim.x) {
        const int t = idx / nzd;
        const int k = idx - t * nzd;
        cplx v = make_double2(0.0, 0.0);
        if (k <= nz)
            v = src[(size_t)t * nzt + nz + k];            // V(iy,0:nz)      -> rows 1..nz+1
        else if (k >= nzd - nz)
            v = src[(size_t)t * nzt + (k - (nzd - nz))];  // V(iy,-nz:-1)    -> rows nzd-nz+1..nzd
        smem[(size_t)t * line_stride + CHB_PAD(k)] = v;
    }
    fft_lines<+1, true>(smem, line_stride, nl, pl, W);
    for (int idx = threadIdx.x; idx < nl * nzd; idx += blockDim.x) {
        const int izd = idx / nl;
        const int t = idx - izd * nl;
        const int peer = izd / g.nzB;
        const int izl = izd - peer * g.nzB;
        Aw.p[peer][chb_bufA_index(g.rank, c, np, pli, g.nzB, izl, g.nxB, ixl0 + t, g.twa)] =
            smem[(size_t)t * line_stride + CHB_PAD(__ldg(&rev[izd]))];
    }
}

// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CONV_THREADS)
zbwd_kernel(const cplx* __restrict__ Br, cplx* __restrict__ P, Geometry g, FftPlan pl, const cplx* __restrict__ W,
            const int* __restrict__ rev, int plane0, int np, int tx, int line_stride) {
    CHB_DYN_SMEM(cplx, smem);
    const int ixl0 = blockIdx.x * tx;
    const int pli = blockIdx.y;
    const int c = blockIdx.z;  // product index 0..5
    const int nl = min(tx, g.nxB - ixl0);
    const int nzd = g.nzd, nz = g.nz, nzt = g.nzt;
    (void)plane0;
    for (int idx = threadIdx.x; idx < nl * nzd; idx += blockDim.x) {
        const int izd = idx / nl;
        const int t = idx - izd * nl;
        const int peer = izd / g.nzB;
        const int izl = izd - peer * g.nzB;
        smem[(size_t)t * line_stride + CHB_PAD(izd)] = Br[chb_bufB_index(peer, 6, c, np, pli, g.nzB, izl, g.nxB, ixl0 + t, g.tw)];
    }
    fft_lines<-1, true>(smem, line_stride, nl, pl, W);
    cplx* dst = P + (((size_t)c * np + pli) * g.nxB + ixl0) * nzt;   // the chunk's spectral products [6][np][nxB][2nz+1]
    for (int idx = threadIdx.x; idx < nl * nzt; idx += blockDim.x) {
        const int t = idx / nzt;
        const int izp = idx - t * nzt;                                // iz + nz
        const int k = (izp >= nz) ? (izp - nz) : (nzd - nz + izp);    // izd(iz), dnsdata.f90:156
        dst[(size_t)t * nzt + izp] = smem[(size_t)t * line_stride + CHB_PAD(__ldg(&rev[k]))];
    }
}

// --------------------------------------------------------------------------------------
// One CTA = lx physical z-lines of one plane.  Shared memory: slot(c,l) = (c*lx + l)*line_stride,
// c=0..5; slots 0..2 hold u,v,w (then uu,vv,ww), slots 3..5 hold uv,vw,uw.
__global__ void __launch_bounds__(CONV_THREADS)
xpass_kernel(const cplx* __restrict__ Ar, PeerPtrs Bw, Geometry g, FftPlan pl, const cplx* __restrict__ W,
             const cplx* __restrict__ Wh, const double* __restrict__ dy, DevScalars* sc, int plane0, int np, int lx,
             int line_stride, int compute_cfl) {
    CHB_DYN_SMEM(cplx, smem);
    __shared__ double red[CONV_THREADS / 32];
    const int izl0 = blockIdx.x * lx;
    const int pli = blockIdx.y;
    const int iy = plane0 + pli - 1;
    const int M = g.nxd;
    const int nx = g.nx;
    const int nxB = g.nxB, nzB = g.nzB;
    // ---- load modes 0..nx of u,v,w, zero-pad to M (dnsdata.f90:535) ----
    for (int idx = threadIdx.x; idx < 3 * lx * M; idx += blockDim.x) {
        const int slot = idx / M;  // c*lx + l
        const int k = idx - slot * M;
        const int c = slot / lx;
        const int l = slot - c * lx;
        cplx v = make_double2(0.0, 0.0);
        if (k <= nx) {
            const int q = k / nxB;
            v = Ar[chb_bufA_index(q, c, np, pli, nzB, izl0 + l, nxB, k - q * nxB, g.twa)];
        }
        smem[(size_t)slot * line_stride + CHB_PAD(k)] = v;
    }
    __syncthreads();
    // ---- c2r of logical length 2M as a complex transform of length M:
    //      Z[k] = (X[k] + conj X[M-k]) + i e^{i pi k/M} (X[k] - conj X[M-k]),  X[M] = 0,
    //      imaginary part of X[0] ignored (FFTW c2r semantics) ----
    const int hp = M / 2 + 1;
    for (int idx = threadIdx.x; idx < 3 * lx * hp; idx += blockDim.x) {
        const int slot = idx / hp;
        const int k = idx - slot * hp;
        cplx* x = smem + (size_t)slot * line_stride;
        if (k == 0) {
            const double r = x[0].x;
            x[0] = make_double2(r, r);
        } else if (2 * k == M) {
            cplx a = x[CHB_PAD(k)];
            x[CHB_PAD(k)] = make_double2(2.0 * a.x, -2.0 * a.y);
        } else {
            const cplx a = x[CHB_PAD(k)];
            const cplx b = x[CHB_PAD(M - k)];
            const cplx s = make_double2(a.x + b.x, a.y - b.y);
            const cplx d = make_double2(a.x - b.x, a.y + b.y);
            const cplx t = cmul(__ldg(&Wh[k]), d);
            x[CHB_PAD(k)] = make_double2(s.x - t.y, s.y + t.x);
            x[CHB_PAD(M - k)] = make_double2(s.x + t.y, t.x - s.y);
        }
    }
    fft_lines<+1, true>(smem, line_stride, 3 * lx, pl, W);
    // ---- physical space (digit-reversed order; pointwise work does not care):
    //      element m of a line holds x[2m] (re) and x[2m+1] (im) ----
    const double f = g.factor;
    double cmax = 0.0;
    const bool do_cfl = compute_cfl && iy >= 1 && iy <= g.ny - 1;
    const double rdy = do_cfl ? dy[iy + 1] : 1.0;
    for (int idx = threadIdx.x; idx < lx * M; idx += blockDim.x) {
        const int l = idx / M;
        const int e = CHB_PAD(idx - l * M);
        cplx* pu = smem + (size_t)(0 * lx + l) * line_stride + e;
        cplx* pv = smem + (size_t)(1 * lx + l) * line_stride + e;
        cplx* pw = smem + (size_t)(2 * lx + l) * line_stride + e;
        const cplx u = *pu, v = *pv, w = *pw;
        if (do_cfl) {  // dnsdata.f90:553-555
            const double c0 = fabs(u.x) / g.dx + fabs(v.x) / rdy + fabs(w.x) / g.dz;
            const double c1 = fabs(u.y) / g.dx + fabs(v.y) / rdy + fabs(w.y) / g.dz;
            cmax = fmax(cmax, fmax(c0, c1));
        }
        smem[(size_t)(3 * lx + l) * line_stride + e] = make_double2(u.x * v.x * f, u.y * v.y * f);  // uv  :581
        smem[(size_t)(4 * lx + l) * line_stride + e] = make_double2(v.x * w.x * f, v.y * w.y * f);  // vw  :582
        smem[(size_t)(5 * lx + l) * line_stride + e] = make_double2(u.x * w.x * f, u.y * w.y * f);  // uw  :583
        *pu = make_double2(u.x * u.x * f, u.y * u.y * f);                                           // :584
        *pv = make_double2(v.x * v.x * f, v.y * v.y * f);
        *pw = make_double2(w.x * w.x * f, w.y * w.y * f);
    }
    if (compute_cfl) {  // block-uniform branch
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cmax;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = red[0];
            for (int i = 1; i < CONV_THREADS / 32; ++i) m = fmax(m, red[i]);
            if (m > 0.0) atomicMax(&sc->cfl_bits, (unsigned long long)__double_as_longlong(m));
        }
    }
    fft_lines<-1, false>(smem, line_stride, 6 * lx, pl, W);
    // ---- r2c post-processing, keep modes 0..nx (x-dealiasing, mpi_transpose.f90:103):
    //      X[k] = (Z[k] + conj Z[M-k])/2 - (i/2) e^{-i pi k/M} (Z[k] - conj Z[M-k]) ----
    const int nxp = nx + 1;
    for (int idx = threadIdx.x; idx < 6 * lx * nxp; idx += blockDim.x) {
        const int slot = idx / nxp;
        const int k = idx - slot * nxp;
        const int c = slot / lx;
        const int l = slot - c * lx;
        const cplx* x = smem + (size_t)slot * line_stride;
        const cplx z = x[CHB_PAD(k)];
        const cplx zm = x[CHB_PAD(k == 0 ? 0 : M - k)];
        const cplx e = make_double2(0.5 * (z.x + zm.x), 0.5 * (z.y - zm.y));
        const cplx d = make_double2(0.5 * (z.x - zm.x), 0.5 * (z.y + zm.y));  // (Z - conj Zm)/2
        const cplx o = make_double2(d.y, -d.x);                               // -i * d
        cplx w = __ldg(&Wh[k]);
        w.y = -w.y;
        const cplx r = cadd(e, cmul(w, o));
        const int q = k / nxB;
        Bw.p[q][chb_bufB_index(g.rank, 6, c, np, pli, nzB, izl0 + l, nxB, k - q * nxB, g.tw)] = r;
    }
}

// --------------------------------------------------------------------------------------
static int pick_lines(int padded_elems, int want_max, size_t limit_bytes) {
    int t = want_max;
    while (t > 1 && (size_t)t * padded_elems * sizeof(cplx) > limit_bytes) t >>= 1;
    return t;
}

static size_t g_smem_optin = 0;
static size_t smem_optin() {
    if (!g_smem_optin) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        g_smem_optin = (size_t)v;
    }
    return g_smem_optin;
}

static void z_config(chb_handle_s* h, int* tx, int* line_stride, size_t* smem) {
    const int ls = chb_padded_len(h->g.nzd);
    const size_t half = smem_optin() / 2 - 2048;  // two CTAs per SM
    int t = pick_lines(ls, 8, half);
    if (t < 4) t = pick_lines(ls, 4, smem_optin() - 1024);
    if (t > h->g.nxB) t = h->g.nxB;
    *tx = t;
    *line_stride = ls;
    *smem = (size_t)t * ls * sizeof(cplx);
}

void launch_zfwd(chb_handle_s* h, int plane0, int nplanes) {
    if (h->use_fft3 && launch_z3_fwd_or_bwd(h, plane0, nplanes, true)) return;
    int tx, ls;
    size_t smem;
    z_config(h, &tx, &ls, &smem);
    cudaFuncSetAttribute(zfwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((h->g.nxB + tx - 1) / tx, nplanes, 3);
    ScopedKernelTimer tm(h, "zfwd", h->cstream);
    CHB_LAUNCH(grid, CONV_THREADS, smem, h->cstream, zfwd_kernel)(h->V, h->Aw, h->g, h->plan_z, h->Wz, h->rev_z, plane0,
                                                         h->chunk_planes, tx, ls);
    h->launches++;
}

void launch_zbwd(chb_handle_s* h, int plane0, int nplanes) {
    if (h->use_fft3 && launch_z3_fwd_or_bwd(h, plane0, nplanes, false)) return;
    int tx, ls;
    size_t smem;
    z_config(h, &tx, &ls, &smem);
    cudaFuncSetAttribute(zbwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((h->g.nxB + tx - 1) / tx, nplanes, 6);
    ScopedKernelTimer tm(h, "zbwd", h->cstream);
    CHB_LAUNCH(grid, CONV_THREADS, smem, h->cstream, zbwd_kernel)(h->Br, h->Pc, h->g, h->plan_z, h->Wz, h->rev_z, plane0,
                                                         h->chunk_planes, tx, ls);
    h->launches++;
}

void launch_xpass(chb_handle_s* h, int plane0, int nplanes, int compute_cfl) {
    if (h->use_fft3 && launch_x3_pass(h, plane0, nplanes, compute_cfl)) return;
    const int ls = chb_padded_len(h->g.nxd);
    // lines per CTA: keep >= ~2 CTAs per SM when possible, and at least ~1.5k butterflies of work
    int lx = 1;
    while (lx < 8 && (h->g.nzB % (lx * 2) == 0) && (size_t)6 * (lx * 2) * ls * sizeof(cplx) <= 48 * 1024) lx *= 2;
    const size_t smem = (size_t)6 * lx * ls * sizeof(cplx);
    cudaFuncSetAttribute(xpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(h->g.nzB / lx, nplanes);
    ScopedKernelTimer tm(h, "xpass", h->cstream);
    CHB_LAUNCH(grid, CONV_THREADS, smem, h->cstream, xpass_kernel)(h->Ar, h->Bw, h->g, h->plan_x, h->Wx, h->Wh, h->t_dy, h->sc,
                                                          plane0, h->chunk_planes, lx, ls, compute_cfl);
    h->launches++;
}

// --------------------------------------------------------------------------------------
// standalone batched FFT used by the FFT parity tests: sign=+-1 runs DIF (natural in,
// permuted readout), sign=+-2 runs DIT (permuted load, natural out); same sign convention.
template <int S, bool DIF>
__global__ void __launch_bounds__(CONV_THREADS)
test_fft_kernel(cplx* data, FftPlan pl, const cplx* __restrict__ W, const int* __restrict__ rev, int nlines, int lpb,
                int line_stride) {
    CHB_DYN_SMEM(cplx, smem);
    const int l0 = blockIdx.x * lpb;
    const int nl = min(lpb, nlines - l0);
    const int n = pl.n;
    for (int idx = threadIdx.x; idx < nl * n; idx += blockDim.x) {
        const int t = idx / n, k = idx - t * n;
        const int pos = DIF ? k : rev[k];
        smem[(size_t)t * line_stride + CHB_PAD(pos)] = data[(size_t)(l0 + t) * n + k];
    }
    fft_lines<S, DIF>(smem, line_stride, nl, pl, W);
    for (int idx = threadIdx.x; idx < nl * n; idx += blockDim.x) {
        const int t = idx / n, k = idx - t * n;
        const int pos = DIF ? rev[k] : k;
        data[(size_t)(l0 + t) * n + k] = smem[(size_t)t * line_stride + CHB_PAD(pos)];
    }
}

int launch_test_fft(const FftPlan& pl, const cplx* W, const int* rev, cplx* data, int nlines, int sign) {
    const int ls = chb_padded_len(pl.n);
    int lpb = pick_lines(ls, 4, 96 * 1024);
    const size_t smem = (size_t)lpb * ls * sizeof(cplx);
    const int grid = (nlines + lpb - 1) / lpb;
#define RUN(S, DIF)                                                                                          \
    cudaFuncSetAttribute(test_fft_kernel<S, DIF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    CHB_LAUNCH(grid, CONV_THREADS, smem, 0, test_fft_kernel<S, DIF>)(data, pl, W, rev, nlines, lpb, ls)
    if (sign == 1) { RUN(+1, true); }
    else if (sign == -1) { RUN(-1, true); }
    else if (sign == 2) { RUN(+1, false); }
    else if (sign == -2) { RUN(-1, false); }
    else { chb_set_error("chb_test_fft_lines: sign must be +-1 (DIF) or +-2 (DIT)"); return 1; }
#undef RUN
    if (cudaGetLastError() != cudaSuccess) { chb_set_error("test_fft launch failed"); return 1; }
    return 0;
}
