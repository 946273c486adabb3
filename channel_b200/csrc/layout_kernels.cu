// layout_kernels.cu - host-layout <-> device-layout transposition.
//
// Host (Fortran) layout: V(-1:ny+1,-nz:nz,nx0:nxN,1:3), iy fastest (dnsdata.f90:132), i.e. C order
// [c][ixl][izp][iyp].  Device layout: [c][iyp][ixl][izp] (the wavenumber index contiguous).  Per
// component this is the 2-D transpose of a [ncols][nyp] matrix, ncols = nxB*(2nz+1).
#include "chb_internal.h"

#define TILE 32

// An x-slab [ix0, ix0 + nix) of one component: `cols` is the slab in host / file order, a contiguous [ncols][nyp] matrix
// (ncols = nix*(2nz+1)); `planes` points at column ix0*(2nz+1) of the component's device array, row length ld = M.
// cols -> planes
__global__ void transpose_cols_to_planes(const cplx* __restrict__ cols, cplx* __restrict__ planes, long long ncols,
                                         int nyp, long long ld) {
    __shared__ cplx tile[TILE][TILE + 1];
    const long long c0 = (long long)blockIdx.x * TILE;
    const int y0 = blockIdx.y * TILE;
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const long long c = c0 + r;
        const int y = y0 + threadIdx.x;
        if (c < ncols && y < nyp) tile[r][threadIdx.x] = cols[c * nyp + y];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const int y = y0 + r;
        const long long c = c0 + threadIdx.x;
        if (c < ncols && y < nyp) planes[(size_t)y * ld + c] = tile[threadIdx.x][r];
    }
}

// planes -> cols
__global__ void transpose_planes_to_cols(const cplx* __restrict__ planes, cplx* __restrict__ cols, long long ncols,
                                         int nyp, long long ld) {
    __shared__ cplx tile[TILE][TILE + 1];
    const long long c0 = (long long)blockIdx.x * TILE;
    const int y0 = blockIdx.y * TILE;
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const int y = y0 + r;
        const long long c = c0 + threadIdx.x;
        if (c < ncols && y < nyp) tile[r][threadIdx.x] = planes[(size_t)y * ld + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TILE; r += blockDim.y) {
        const long long c = c0 + r;
        const int y = y0 + threadIdx.x;
        if (c < ncols && y < nyp) cols[c * nyp + y] = tile[threadIdx.x][r];
    }
}

void launch_fortran_to_planes(chb_handle_s* h, const cplx* cols, cplx* planes, int ix0, int nix, cudaStream_t st) {
    const long long ncols = (long long)nix * h->g.nzt;
    dim3 grid((unsigned)((ncols + TILE - 1) / TILE), (h->g.nyp + TILE - 1) / TILE), block(TILE, 8);
    CHB_LAUNCH(grid, block, 0, st, transpose_cols_to_planes)(cols, planes + (size_t)ix0 * h->g.nzt, ncols, h->g.nyp, (long long)h->g.M);
    h->launches++;
}
void launch_planes_to_fortran(chb_handle_s* h, const cplx* planes, cplx* cols, int ix0, int nix, cudaStream_t st) {
    const long long ncols = (long long)nix * h->g.nzt;
    dim3 grid((unsigned)((ncols + TILE - 1) / TILE), (h->g.nyp + TILE - 1) / TILE), block(TILE, 8);
    CHB_LAUNCH(grid, block, 0, st, transpose_planes_to_cols)(planes + (size_t)ix0 * h->g.nzt, cols, ncols, h->g.nyp, (long long)h->g.M);
    h->launches++;
}
