// fft_emul.cpp - test infrastructure: the x-pass kernel of the nonlinear term (xpass3_kernels.cu) compiled with
// g++ and run on CPU threads (cta_emul.hpp: one OS thread per CUDA thread, __syncthreads = barrier).  Lets the
// CPU suite check the kernel's indexing / twiddles / shared-memory layout against numpy for a few z-lines of the
// large transform sizes, and lets a new kernel variant be proven correct before it ever sees a GPU.
#include "cta_emul.hpp"

#include <cmath>

#define CHB_HOST_EMUL 1
#include "../../channel_b200/csrc/xpass3_kernels.cu"

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
                     compute_cfl);
}

template <class G, int MINB>
static void run_x5(const cplx* Ar, cplx* Br, const Geometry& g, const cplx* W, const cplx* Wh, const double* dy, DevScalars* sc,
                   int np, int compute_cfl) {
    PeerPtrs Bw;
    memset(&Bw, 0, sizeof(Bw));
    Bw.p[0] = Br;
    constexpr int T = 2 * (G::N / G::C);
    cta_emul::launch(xpass5_kernel<G, MINB, false>, dim3(g.nzB, np, 1), T, Ar, Bw, g, W, Wh, dy, sc, 0, np, compute_cfl);
}

extern "C" {

// One chunk of `np` planes, single rank: Ar = velocities after the z pass, row-major [3][np][nzB][nx+1];
// Br = products for the backward z pass in the tiled layout of transpose_index.h (tile width 2^tw),
// [6][np][(nx+1) >> tw][nzB][1 << tw]; dy[ny+3]; cfl_out = max of the CFL expression (dnsdata.f90:552-556).
// Planes are iy = -1 .. np-2 (plane0 = 0).  variant 0: xpass4 (one thread per innermost butterfly position),
// 1: xpass5 (two threads per position; nxd = 1536).  Returns 2 if no such kernel exists for nxd.
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
        if (nxd != 1536) return 2;
        run_x5<Fft3<1536, 12, 16, 8>, 1>(A, B, g, Wc, Whc, dy, &sc, np, compute_cfl);
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

}  // extern "C"
