// convvel.cu - the convection-velocity diagnostic of the reference (#ifdef convvel; dnsdata.f90:84-89, 141-143,
// 515-531, 546-549, 858-860, 908-913 and save_convvel_file :792-816).
//
// In the first convolutions sweep after every outstats the reference compares, for every velocity component, plane,
// x-mode ix > 0 and physical z point, the z-transformed velocity with the one stored a time step earlier:
//   dtu = (u - uold)/deltat,  ust = (u + uold)/2,  cu = Im(conj(ust) dtu) / (ix alfa0 |ust|^2),  uconv += cu
// and keeps u for the next time.  Here the z-transformed velocity of a chunk of planes sits in the work buffer of the
// pencil transpose right after zfwd, so the diagnostic is one element-wise kernel per chunk on it; Voldz and uconv
// live in HBM as [3][ny+3][nzd][nx+1] (4.5 + 2.25 complex-equivalents per point: a research option, off by default
// like the reference's commented-out #define).  One GPU only in this version: with several, the buffer a rank sees
// after zTOx holds other z-lines than the reference's x-slab owner does, and the file writer would have to gather.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/channel_b200.h"
#include "chb_internal.h"

__global__ void convvel_kernel(const cplx* __restrict__ Ar, cplx* __restrict__ Vold, double* __restrict__ uconv, Geometry g,
                               int plane0, int np_alloc, double deltat, int accumulate) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (z row, x) of one plane
    const long long per_plane = (long long)g.nzB * g.nxB;
    if (e >= per_plane) return;
    const int pli = blockIdx.y, comp = blockIdx.z;
    const int ix0 = g.nx0 + (int)(e % g.nxB);
    const cplx u = Ar[((size_t)comp * np_alloc + pli) * per_plane + e];
    const size_t p = ((size_t)comp * g.nyp + plane0 + pli) * per_plane + e;
    if (accumulate && ix0 > 0) {                                             // dnsdata.f90:517-528
        const cplx o = Vold[p];
        const double dx = (u.x - o.x) / deltat, dy = (u.y - o.y) / deltat;   // dtu
        const double sx = 0.5 * (u.x + o.x), sy = 0.5 * (u.y + o.y);         // ust
        uconv[p] += (sx * dy - sy * dx) / ((ix0 * g.alfa0) * (sx * sx + sy * sy));
    }
    Vold[p] = u;                                                             // :530
}

#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)
// called by convolutions_all after zfwd (+ zTOx) of a chunk when this sweep computes the diagnostic
void launch_convvel(chb_handle_s* h, int plane0, int nplanes, double deltat) {
    const Geometry& g = h->g;
    const long long per_plane = (long long)g.nzB * g.nxB;
    dim3 grid((unsigned)((per_plane + 255) / 256), nplanes, 3);
    ScopedKernelTimer tm(h, "convvel", h->cstream);
    CHB_LAUNCH(grid, 256, 0, h->cstream, convvel_kernel)(h->Ar, h->cv_Vold, h->cv_uconv, g, plane0, h->chunk_planes, deltat,
                                                         h->cv_cnt > -1 ? 1 : 0);
    h->launches++;
}

#define CV_REQUIRE(cond, msg)   \
    do {                        \
        if (!(cond)) {          \
            chb_set_error(msg); \
            return 2;           \
        }                       \
    } while (0)

extern "C" int chb_set_convvel(chb_handle h, int enable) {
    CV_REQUIRE(h, "chb_set_convvel: null handle");
    CV_REQUIRE(!enable || h->g.nranks == 1, "chb_set_convvel: the convection-velocity diagnostic runs on one GPU only in this version");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (h->cv_Vold) { cudaFree(h->cv_Vold); h->cv_Vold = nullptr; }
    if (h->cv_uconv) { cudaFree(h->cv_uconv); h->cv_uconv = nullptr; }
    h->cv_enabled = 0;
    if (!enable) return 0;
    const size_t n = (size_t)3 * h->g.nyp * h->g.nzd * h->g.nxB;
    CHB_CUDA_OK(cudaMalloc((void**)&h->cv_Vold, n * sizeof(cplx)));            // Voldz = 0; uconv = 0   dnsdata.f90:142
    CHB_CUDA_OK(cudaMalloc((void**)&h->cv_uconv, n * sizeof(double)));
    CHB_CUDA_OK(cudaMemset(h->cv_Vold, 0, n * sizeof(cplx)));
    CHB_CUDA_OK(cudaMemset(h->cv_uconv, 0, n * sizeof(double)));
    h->dev_bytes += n * (sizeof(cplx) + sizeof(double));
    h->cv_enabled = 1;
    h->cv_cnt = -1;            // convvel_cnt = -1, compute_convvel = .FALSE.   dnsdata.f90:87-88
    h->cv_compute = 0;
    return 0;
}

// uconv in the order save_convvel_file writes it, [iV][iy+1][ix][iz_d] (dnsdata.f90:803-808), not divided by the count
static int fetch_uconv(chb_handle h, std::vector<double>* out) {
    const Geometry& g = h->g;
    const size_t n = (size_t)3 * g.nyp * g.nzd * g.nxB;
    std::vector<double> dev(n);
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaMemcpy(dev.data(), h->cv_uconv, n * sizeof(double), cudaMemcpyDeviceToHost));
    out->resize(n);
    const size_t nzd = g.nzd, nxB = g.nxB;
    for (size_t cp = 0; cp < (size_t)3 * g.nyp; ++cp)          // device order [iV][iy+1][iz_d][ix]
        for (size_t z = 0; z < nzd; ++z)
            for (size_t x = 0; x < nxB; ++x) (*out)[(cp * nxB + x) * nzd + z] = dev[(cp * nzd + z) * nxB + x];
    return 0;
}

extern "C" int chb_get_convvel(chb_handle h, double* uconv_host, long long* count) {
    CV_REQUIRE(h && h->cv_enabled, "chb_get_convvel: the diagnostic is not enabled (chb_set_convvel)");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    if (count) *count = h->cv_cnt;
    if (uconv_host) {
        std::vector<double> v;
        if (fetch_uconv(h, &v)) return 1;
        memcpy(uconv_host, v.data(), v.size() * sizeof(double));
    }
    return 0;
}

// what outstats does at the dt_field cadence (dnsdata.f90:908-913): uconv/convvel_cnt to Convvel.cart.<n>.out, then reset
extern "C" int chb_save_convvel_file(chb_handle h, const char* filename) {
    CV_REQUIRE(h && filename && h->cv_enabled, "chb_save_convvel_file: the diagnostic is not enabled (chb_set_convvel)");
    CHB_CUDA_OK(cudaSetDevice(h->device));
    std::vector<double> v;
    if (fetch_uconv(h, &v)) return 1;
    const double c = (double)h->cv_cnt;
    for (double& x : v) x = x / c;
    FILE* f = fopen(filename, "wb");
    if (!f) { chb_set_error(std::string("chb_save_convvel_file: cannot open ") + filename + ": " + strerror(errno)); return 4; }
    const size_t w = fwrite(v.data(), sizeof(double), v.size(), f);
    if (fclose(f) != 0 || w != v.size()) { chb_set_error("chb_save_convvel_file: short write"); return 4; }
    CHB_CUDA_OK(cudaMemset(h->cv_uconv, 0, v.size() * sizeof(double)));
    h->cv_cnt = 0;
    return 0;
}
#endif
