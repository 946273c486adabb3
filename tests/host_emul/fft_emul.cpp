// fft_emul.cpp - test infrastructure: the x-pass kernel of the nonlinear term (xpass3_kernels.cu) compiled with
// g++ and run on CPU threads (cta_emul.hpp: one OS thread per CUDA thread, __syncthreads = barrier).  Lets the
// CPU suite check the kernel's indexing / twiddles / shared-memory layout against numpy for a few z-lines of the
// large transform sizes, and lets a new kernel variant be proven correct before it ever sees a GPU.
#include "cta_emul.hpp"

#include <cmath>

#define CHB_HOST_EMUL 1
#include "../../channel_b200/csrc/xpass3_kernels.cu"
#include "../../channel_b200/csrc/zpass3_kernels.cu"

static std::vector<double> twiddle_table(int count, int denom) {   // as chb_api.cu: exp(+2 pi i e / denom)
    std::vector<double> w(2 * (size_t)count);
    for (int e = 0; e < count; ++e) {
        const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / (long double)denom;
        w[2 * e] = (double)cosl(a);
        w[2 * e + 1] = (double)sinl(a);
    }
    return w;
}

template <class G, int LPC, int MINB>
static void run_x4(const cplx* Ar, cplx* Br, const Geometry& g, const cplx* W, const cplx* Wh, const double* dy, DevScalars* sc,
                   int np, int compute_cfl) {
    PeerPtrs Bw;
    memset(&Bw, 0, sizeof(Bw));
    Bw.p[0] = Br;
    constexpr int T = G::N / G::C;
    cta_emul::launch(xpass4_kernel<G, LPC, MINB, false>, dim3(g.nzB / LPC, np, 1), LPC * T, Ar, Bw, g, W, Wh, dy, sc, 0, np,
                     compute_cfl, np);
}

// persistent variant (inputs of the next line prefetched into shared memory): two CTAs walk over the nzB * np lines
template <class G, int MINB, int SPLIT>
static void run_xp(const cplx* Ar, cplx* Br, const Geometry& g, const cplx* W, const cplx* Wh, const double* dy, DevScalars* sc,
                   int np, int compute_cfl) {
    PeerPtrs Bw;
    memset(&Bw, 0, sizeof(Bw));
    Bw.p[0] = Br;
    constexpr int T = SPLIT * (G::N / G::C);
    constexpr int LB = G::A * (G::BC + 1);
    if (((size_t)6 * LB + 4 * (g.nx + 1) + G::BC + G::B) * sizeof(cplx) > sizeof(cta_emul::g_dyn_smem)) abort();
    cta_emul::launch(xpass4_kernel<G, 1, MINB, false, SPLIT, true>, dim3(2, 1, 1), T, Ar, Bw, g, W, Wh, dy, sc, 0, np, compute_cfl, np);
}

template <class G, int MINB>
static void run_x5(const cplx* Ar, cplx* Br, const Geometry& g, const cplx* W, const cplx* Wh, const double* dy, DevScalars* sc,
                   int np, int compute_cfl) {
    PeerPtrs Bw;
    memset(&Bw, 0, sizeof(Bw));
    Bw.p[0] = Br;
    constexpr int T = 2 * (G::N / G::C);
    cta_emul::launch(xpass4_kernel<G, 1, MINB, false, 2>, dim3(g.nzB, np, 1), T, Ar, Bw, g, W, Wh, dy, sc, 0, np, compute_cfl, np);
}

template <class G, int LPC, int TPL, int MINB>
static void run_z4(int fwd, const cplx* in, cplx* out, const Geometry& g, const cplx* W, int np) {
    constexpr int BCP = G::BC + 1;
    int LS = G::A * BCP;
    const int want = (LPC == 8) ? 1 : (LPC == 4 ? 2 : 4);       // as launch_z4 (zpass3_kernels.cu)
    while (LS % 8 != want) ++LS;
    if ((size_t)LPC * LS * sizeof(cplx) > sizeof(cta_emul::g_dyn_smem)) abort();
    if (fwd) {
        PeerPtrs Aw;
        memset(&Aw, 0, sizeof(Aw));
        Aw.p[0] = out;
        if (fwd == 2)   // DIRECT: stage A reads global memory
            cta_emul::launch(zfwd4_kernel<G, LPC, TPL, MINB, true>, dim3(g.nxB / LPC, np, 3), LPC * TPL, in, Aw, g, W, 0, np, LS);
        else
            cta_emul::launch(zfwd4_kernel<G, LPC, TPL, MINB, false>, dim3(g.nxB / LPC, np, 3), LPC * TPL, in, Aw, g, W, 0, np, LS);
    } else {
        cta_emul::launch(zbwd4_kernel<G, LPC, TPL, MINB>, dim3(g.nxB / LPC, np, 6), LPC * TPL, in, out, g, W, 0, np, LS);
    }
}

extern "C" {

// One chunk of `np` planes, single rank: Ar = velocities after the z pass, row-major [3][np][nzB][nx+1];
// Br = products for the backward z pass in the tiled layout of transpose_index.h (tile width 2^tw),
// [6][np][(nx+1) >> tw][nzB][1 << tw]; dy[ny+3]; cfl_out = max of the CFL expression (dnsdata.f90:552-556).
// Planes are iy = -1 .. np-2 (plane0 = 0).  variant 0: xpass4 (one thread per innermost butterfly position),
// 1: the split x-pass (two threads per position; nxd = 768, 1536), 2 / 3: the persistent x-pass (prefetched inputs) with one / two threads per position.  Returns 2 if no such kernel exists for nxd.
__attribute__((visibility("default"))) int chb_emul_xpass(int nx, int ny, int nzB, int np, int nxd, int nzd, double alfa0,
                                                          double beta0, int tw, const double* Ar, double* Br,
                                                          const double* dy, int compute_cfl, double* cfl_out, int variant) {
    Geometry g;
    memset(&g, 0, sizeof(g));
    g.nx = nx; g.ny = ny; g.nxd = nxd; g.nzd = nzd;
    g.nyp = ny + 3;
    g.rank = 0; g.nranks = 1;
    g.nx0 = 0; g.nxN = nx; g.nxB = nx + 1;
    g.nz0 = 0; g.nzN = nzB - 1; g.nzB = nzB;
    const double PI = 3.1415926535897932384626433832795028841971;
    g.dx = PI / (alfa0 * nxd); g.dz = 2.0 * PI / (beta0 * nzd); g.factor = 1.0 / (2.0 * nxd * nzd);
    g.tw = tw; g.twa = -1;
    std::vector<double> W = twiddle_table(nxd, nxd), Wh = twiddle_table(nxd, 2 * nxd);
    DevScalars sc;
    memset(&sc, 0, sizeof(sc));
    const cplx* A = reinterpret_cast<const cplx*>(Ar);
    cplx* B = reinterpret_cast<cplx*>(Br);
    const cplx* Wc = reinterpret_cast<const cplx*>(W.data());
    const cplx* Whc = reinterpret_cast<const cplx*>(Wh.data());
    if (variant == 1) {
        if (nxd == 1536) run_x5<Fft3<1536, 12, 16, 8>, 1>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
        else if (nxd == 768) run_x5<Fft3<768, 12, 16, 4>, 2>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
        else return 2;
    } else if (variant == 2 || variant == 3) {   // persistent, one / two threads per position
        if (nxd == 1536 && variant == 2) run_xp<Fft3<1536, 12, 16, 8>, 1, 1>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
        else if (nxd == 1536) run_xp<Fft3<1536, 12, 16, 8>, 1, 2>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
        else if (nxd == 768 && variant == 2) run_xp<Fft3<768, 12, 16, 4>, 2, 1>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
        else if (nxd == 768) run_xp<Fft3<768, 12, 16, 4>, 2, 2>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
        else return 2;
    } else
    switch (nxd) {
        case 384: run_x4<Fft3<384, 12, 8, 4>, 1, 6>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl); break;
        case 768: run_x4<Fft3<768, 12, 16, 4>, 1, 3>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl); break;
        case 1536: run_x4<Fft3<1536, 12, 16, 8>, 1, 1>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl); break;
        default: return 2;
    }
    double c;
    memcpy(&c, &sc.cfl_bits, sizeof(double));
    if (cfl_out) *cfl_out = c;
    return 0;
}

// z passes of one chunk of `np` planes (plane0 = 0, so ny+3 must equal np), single rank, nzB = nzd.
//   fwd = 1 (TMA-staged stage A) or 2 (direct global loads): in = V [3][np][nxB][2nz+1] -> out = velocity work buffer (transpose_index.h, tile width 2^twa or
//             row-major [3][np][nzd][nxB] for twa < 0)
//   fwd == 0: in = products work buffer (tile width 2^tw) -> out = P [6][np][nxB][2nz+1]
// lpc = lines per CTA (the variants chb_create selects from).  Returns 2 if that kernel does not exist.
__attribute__((visibility("default"))) int chb_emul_zpass(int fwd, int nxB, int nz, int nzd, int np, int lpc, int tw, int twa,
                                                          const double* in, double* out) {
    Geometry g;
    memset(&g, 0, sizeof(g));
    g.nz = nz; g.nzd = nzd; g.nzt = 2 * nz + 1;
    g.ny = np - 3; g.nyp = np;
    g.rank = 0; g.nranks = 1;
    g.nx = nxB - 1; g.nx0 = 0; g.nxN = nxB - 1; g.nxB = nxB;
    g.nz0 = 0; g.nzN = nzd - 1; g.nzB = nzd;
    g.M = (long long)nxB * g.nzt;
    g.tw = tw; g.twa = twa;
    if (nxB % (lpc & 1023) != 0) return 2;
    std::vector<double> W = twiddle_table(nzd, nzd);
    const cplx* Wc = reinterpret_cast<const cplx*>(W.data());
    const cplx* I = reinterpret_cast<const cplx*>(in);
    cplx* O = reinterpret_cast<cplx*>(out);
    const int f = fwd;
    switch (nzd * 16 + lpc) {     // as launch_z3_fwd_or_bwd
        case 768 * 16 + 2: run_z4<Fft3<768, 12, 8, 8>, 2, 64, 8>(f, I, O, g, Wc, np); break;
        case 768 * 16 + 4: run_z4<Fft3<768, 12, 8, 8>, 4, 64, 4>(f, I, O, g, Wc, np); break;
        case 768 * 16 + 8: run_z4<Fft3<768, 12, 8, 8>, 8, 32, 2>(f, I, O, g, Wc, np); break;
        case 1536 * 16 + 2: run_z4<Fft3<1536, 12, 16, 8>, 2, 64, 4>(f, I, O, g, Wc, np); break;
        case 1536 * 16 + 4: run_z4<Fft3<1536, 12, 16, 8>, 4, 64, 2>(f, I, O, g, Wc, np); break;
        case 1536 * 16 + 8: run_z4<Fft3<1536, 12, 16, 8>, 8, 32, 1>(f, I, O, g, Wc, np); break;
        case 3072 * 16 + 2: run_z4<Fft3<3072, 12, 16, 16>, 2, 64, 2>(f, I, O, g, Wc, np); break;
        case 3072 * 16 + 4: run_z4<Fft3<3072, 12, 16, 16>, 4, 64, 1>(f, I, O, g, Wc, np); break;
        case 3072 * 16 + 2 + 1024: run_z4<Fft3<3072, 12, 16, 16>, 2, 128, 2>(f, I, O, g, Wc, np); break;   // lpc = 2 + 1024: TPL = 128
        case 1536 * 16 + 2 + 1024: run_z4<Fft3<1536, 12, 16, 8>, 2, 128, 3>(f, I, O, g, Wc, np); break;
        case 1536 * 16 + 2 + 2048: run_z4<Fft3<1536, 12, 16, 8>, 2, 96, 4>(f, I, O, g, Wc, np); break;      // + 2048: TPL = 96
        default: return 2;
    }
    return 0;
}

// The whole nonlinear term of one plane on P emulated ranks with the direct-store transposes (CHB_P2P=1): rank r owns
// x-modes [r nxB, (r+1) nxB) and z-lines [r nzB, (r+1) nzB); zfwd of every rank stores into the owners' velocity
// buffers, xpass of every rank into the owners' product buffers (PeerPtrs = the other ranks' buffers, as CUDA IPC
// maps them), zbwd reads its own.  nx = 255 (nxd = 384), nz = 255 (nzd = 768), one plane (iy = 1 of ny = 8), the
// default lines-per-CTA (zfwd 4, zbwd 4) and tile widths chb_create would choose.
//   V: [P][3][1][nxB][2nz+1] per-rank slabs -> Pout: [P][6][1][nxB][2nz+1]
__attribute__((visibility("default"))) int chb_emul_convolutions_multi(int P, const double* V, double* Pout) {
    const int nx = 255, nz = 255, nxd = 384, nzd = 768, np = 1;
    if ((nx + 1) % P || nzd % P || P > CHB_MAX_RANKS) return 2;
    const int nxB = (nx + 1) / P, nzB = nzd / P, nzt = 2 * nz + 1;
    const size_t na = (size_t)3 * np * nzd * nxB, nb = (size_t)6 * np * nzd * nxB, nv = (size_t)np * nxB * nzt;
    std::vector<std::vector<cplx>> Ar(P, std::vector<cplx>(na)), Br(P, std::vector<cplx>(nb));
    std::vector<double> Wz = twiddle_table(nzd, nzd), Wx = twiddle_table(nxd, nxd), Wh = twiddle_table(nxd, 2 * nxd);
    std::vector<double> dy(16, 1.0);
    auto geom = [&](int r) {
        Geometry g;
        memset(&g, 0, sizeof(g));
        g.nx = nx; g.ny = 8; g.nz = nz; g.nxd = nxd; g.nzd = nzd; g.nyp = np; g.nzt = nzt;
        g.rank = r; g.nranks = P;
        chb_decompose(nx + 1, nzd, P, r, &g.nx0, &g.nxN, &g.nz0, &g.nzN);
        g.nxB = nxB; g.nzB = nzB; g.M = (long long)nxB * nzt;
        const double PI = 3.1415926535897932384626433832795028841971;
        g.alfa0 = 0.5; g.beta0 = 1.0; g.dx = PI / (0.5 * nxd); g.dz = 2.0 * PI / nzd; g.factor = 1.0 / (2.0 * nxd * nzd);
        g.tw = (nxB % 8 == 0) ? 3 : ((nxB % 4 == 0) ? 2 : 0);     // as chb_create
        g.twa = (P > 1 && nxB % 4 == 0) ? 2 : -1;                  // width of a zfwd CTA (4 lines)
        return g;
    };
    PeerPtrs Aw, Bw;
    memset(&Aw, 0, sizeof(Aw)); memset(&Bw, 0, sizeof(Bw));
    for (int q = 0; q < P; ++q) { Aw.p[q] = Ar[q].data(); Bw.p[q] = Br[q].data(); }
    typedef Fft3<768, 12, 8, 8> GZ;
    typedef Fft3<384, 12, 8, 4> GX;
    constexpr int LPC = 4, TPL = 64;
    int LS = GZ::A * (GZ::BC + 1);
    while (LS % 8 != 2) ++LS;
    DevScalars sc;
    memset(&sc, 0, sizeof(sc));
    for (int r = 0; r < P; ++r)   // z-pad + backward z FFT + zTOx into the owners' buffers
        cta_emul::launch(zfwd4_kernel<GZ, LPC, TPL, 4, false>, dim3(nxB / LPC, np, 3), LPC * TPL,
                         reinterpret_cast<const cplx*>(V) + (size_t)r * 3 * nv, Aw, geom(r), reinterpret_cast<const cplx*>(Wz.data()), 0, np, LS);
    const bool persist = getenv("CHB_EMUL_XPERSIST") != nullptr;   // the persistent x-pass (prefetched inputs), two CTAs per rank
    for (int r = 0; r < P; ++r) { // x pass on the z-lines of rank r, xTOz into the owners' buffers
        const cplx* a = (const cplx*)Ar[r].data();
        const cplx* wx = reinterpret_cast<const cplx*>(Wx.data());
        const cplx* wh = reinterpret_cast<const cplx*>(Wh.data());
        const double* dyp = (const double*)dy.data();
        if (P > 1 && persist) cta_emul::launch(xpass4_kernel<GX, 1, 6, true, 1, true>, dim3(2, 1, 1), GX::N / GX::C, a, Bw, geom(r), wx, wh, dyp, &sc, 0, np, 0, np);
        else if (persist) cta_emul::launch(xpass4_kernel<GX, 1, 6, false, 1, true>, dim3(2, 1, 1), GX::N / GX::C, a, Bw, geom(r), wx, wh, dyp, &sc, 0, np, 0, np);
        else if (P > 1) cta_emul::launch(xpass4_kernel<GX, 1, 6, true>, dim3(nzB, np, 1), GX::N / GX::C, a, Bw, geom(r), wx, wh, dyp, &sc, 0, np, 0, np);
        else cta_emul::launch(xpass4_kernel<GX, 1, 6, false>, dim3(nzB, np, 1), GX::N / GX::C, a, Bw, geom(r), wx, wh, dyp, &sc, 0, np, 0, np);
    }
    for (int r = 0; r < P; ++r)   // forward z FFT + truncation
        cta_emul::launch(zbwd4_kernel<GZ, LPC, TPL, 4>, dim3(nxB / LPC, np, 6), LPC * TPL, (const cplx*)Br[r].data(),
                         reinterpret_cast<cplx*>(Pout) + (size_t)r * 6 * nv, geom(r), reinterpret_cast<const cplx*>(Wz.data()), 0, np, LS);
    return 0;
}

}  // extern "C"
