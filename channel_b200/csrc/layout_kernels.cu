// layout_kernels.cu - host-layout <-> device-layout transposition.
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
    CHB_LAUNCH(grid, block, 0, h->stream, transpose_cols_to_planes)(src, dst, ncols, h->g.nyp);
    h->launches++;
}
void launch_planes_to_fortran(chb_handle_s* h, const cplx* src, cplx* dst, int, int, int) {
    const long long ncols = h->g.M;
    dim3 grid((unsigned)((ncols + TILE - 1) / TILE), (h->g.nyp + TILE - 1) / TILE), block(TILE, 8);
    CHB_LAUNCH(grid, block, 0, h->stream, transpose_planes_to_cols)(src, dst, ncols, h->g.nyp);
    h->launches++;
}
