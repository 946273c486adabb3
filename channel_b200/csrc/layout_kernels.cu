// layout_kernels.cu - host-layout <-> device-layout transposition and the body-force kernels.
//
// Host (Fortran) layout: V(-1:ny+1,-nz:nz,nx0:nxN,1:3), iy fastest (dnsdata.f90:132), i.e. C order
// [c][ixl][izp][iyp].  Device layout: [c][iyp][ixl][izp] (the wavenumber index contiguous).  Per
// component this is the 2-D transpose of a [ncols][nyp] matrix, ncols = nxB*(2nz+1).
#include "chb_internal.h"

#define TILE 32

// src [ncols][nyp] -> dst [nyp][ncols]
__global__ void transpose_cols_to_planes(const cplx* __restrict__ src, cplx* __restrict__ dst, long long ncols,
                                         int nyp) {
    __shared__ cplx tile[TILE][TILE + 1];
    const long long c0 = (long long)blockIdx.x * TILE;
    const int y0 = blockIdx.y * TILE;
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const long long c = c0 + r;
        const int y = y0 + threadIdx.x;
        if (c < ncols && y < nyp) tile[r][threadIdx.x] = src[c * nyp + y];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const int y = y0 + r;
        const long long c = c0 + threadIdx.x;
        if (c < ncols && y < nyp) dst[(size_t)y * ncols + c] = tile[threadIdx.x][r];
    }
}

// src [nyp][ncols] -> dst [ncols][nyp]
__global__ void transpose_planes_to_cols(const cplx* __restrict__ src, cplx* __restrict__ dst, long long ncols,
                                         int nyp) {
    __shared__ cplx tile[TILE][TILE + 1];
    const long long c0 = (long long)blockIdx.x * TILE;
    const int y0 = blockIdx.y * TILE;
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const int y = y0 + r;
        const long long c = c0 + threadIdx.x;
        if (c < ncols && y < nyp) tile[r][threadIdx.x] = src[(size_t)y * ncols + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const long long c = c0 + r;
        const int y = y0 + threadIdx.x;
        if (c < ncols && y < nyp) dst[c * nyp + y] = tile[threadIdx.x][r];
    }
}

void launch_fortran_to_planes(chb_handle_s* h, const cplx* src, cplx* dst, int, int, int) {
    const long long ncols = h->g.M;
    dim3 grid((unsigned)((ncols + TILE - 1) / TILE), (h->g.nyp + TILE - 1) / TILE), block(TILE, 8);
    transpose_cols_to_planes<<<grid, block, 0, h->stream>>>(src, dst, ncols, h->g.nyp);
    h->launches++;
}
void launch_planes_to_fortran(chb_handle_s* h, const cplx* src, cplx* dst, int, int, int) {
    const long long ncols = h->g.M;
    dim3 grid((unsigned)((ncols + TILE - 1) / TILE), (h->g.nyp + TILE - 1) / TILE), block(TILE, 8);
    transpose_planes_to_cols<<<grid, block, 0, h->stream>>>(src, dst, ncols, h->g.nyp);
    h->launches++;
}

// F_r = sum_c A[r][c] * V_c inside mask_y(iy)*mask_z(iz) (body_forces/*/*.inc set_body_force)
__global__ void body_force_kernel(const cplx* __restrict__ V, cplx* __restrict__ F, Geometry g, BodyForce bf) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int iyp = blockIdx.y;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    if (bf.mask_y[iyp] == 0.0 || bf.mask_z[izp] == 0.0) return;
    if (bf.exclude_mean && g.nx0 + ixl == 0 && izp == g.nz) return;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    const size_t off = (size_t)iyp * plane + m;
    const cplx u = V[off], v = V[comp + off], w = V[2 * comp + off];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        cplx f;
        f.x = bf.A[r * 3 + 0] * u.x + bf.A[r * 3 + 1] * v.x + bf.A[r * 3 + 2] * w.x;
        f.y = bf.A[r * 3 + 0] * u.y + bf.A[r * 3 + 1] * v.y + bf.A[r * 3 + 2] * w.y;
        F[r * comp + off] = f;
    }
}

// ghost extension of F at the start of buildrhs (dnsdata.f90:616-629)
__global__ void force_ghost_kernel(cplx* __restrict__ F, Geometry g, DevTables tab) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    const int ny = g.ny;
    for (int c = 0; c < 3; ++c) {
        cplx* f = F + c * comp + m;
        // F(-1:0)=0 ; F(-1) = -D4(F)|iy=1 / d4(1)(-2)
        {
            const double* d4 = tab.d4 + (1 + 1) * 5;
            cplx s = make_double2(0, 0);
            for (int j = 2; j < 5; ++j) {  // F(-1)=F(0)=0 -> only nodes 1..3 contribute
                const cplx a = f[(size_t)(j) * plane];  // node iy = 1-2+j -> index j
                s.x += d4[j] * a.x;
                s.y += d4[j] * a.y;
            }
            f[(size_t)1 * plane] = make_double2(0, 0);
            f[(size_t)0 * plane] = make_double2(-s.x / d4[0], -s.y / d4[0]);
        }
        {
            const double* d4 = tab.d4 + (ny - 1 + 1) * 5;
            cplx s = make_double2(0, 0);
            for (int j = 0; j < 3; ++j) {  // nodes ny-3..ny-1
                const cplx a = f[(size_t)(ny - 3 + j + 1) * plane];
                s.x += d4[j] * a.x;
                s.y += d4[j] * a.y;
            }
            f[(size_t)(ny + 1) * plane] = make_double2(0, 0);
            f[(size_t)(ny + 2) * plane] = make_double2(-s.x / d4[4], -s.y / d4[4]);
        }
    }
}

void launch_body_force(chb_handle_s* h) {
    const Geometry& g = h->g;
    dim3 grid((unsigned)((g.M + 255) / 256), g.nyp);
    body_force_kernel<<<grid, 256, 0, h->stream>>>(h->V, h->F, g, h->bf);
    h->launches++;
}
void launch_force_ghosts(chb_handle_s* h) {
    const Geometry& g = h->g;
    force_ghost_kernel<<<(unsigned)((g.M + 255) / 256), 256, 0, h->stream>>>(h->F, g, h->tab);
    h->launches++;
}
