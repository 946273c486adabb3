// transpose_index.h - layout of the pencil-transpose work buffers, shared by the kernels
// (conv_kernels.cu), the all-to-all (transpose.cu) and the host-side tests.
//
// One contiguous block per peer rank so that zTOx / xTOz (mpi_transpose.f90:50-117) are plain
// block exchanges:  buf[peer][comp][plane][izl][ixl],  izl in [0,nzB), ixl in [0,nxB).
// On the z side (written by the z-pass, rank r owns x-modes nx0..nxN) peer = owner of the
// physical z-line = iz_d / nzB; after the exchange, on the x side (rank q owns z-lines
// nz0..nzN) peer = owner of the x-mode = ix / nxB.
#pragma once
#include <stddef.h>
#ifdef __CUDACC__
#define CHB_HD __host__ __device__ __forceinline__
#else
#define CHB_HD inline
#endif

CHB_HD size_t chb_buf_index(int peer, int ncomp, int comp, int np, int pl, int nzB, int izl, int nxB, int ixl) {
    return ((((size_t)peer * ncomp + comp) * np + pl) * nzB + izl) * (size_t)nxB + ixl;
}

// The products buffer (x-pass -> z-pass, xTOz) is tiled: 2^tw consecutive x-modes form the
// innermost index, then the z row, so that the LPC = 2^tw lines one z-pass CTA transforms are one
// contiguous nzB * 2^tw * 16-byte block:  buf[peer][comp][plane][ixl >> tw][izl][ixl & (2^tw-1)].
// tw = 0 is the fully transposed layout [ixl][izl].
CHB_HD size_t chb_bufB_index(int peer, int ncomp, int comp, int np, int pl, int nzB, int izl, int nxB, int ixl, int tw) {
    return (((((size_t)peer * ncomp + comp) * np + pl) * (size_t)(nxB >> tw) + (ixl >> tw)) * nzB + izl) * ((size_t)1 << tw) +
           (ixl & ((1 << tw) - 1));
}

// The velocity buffer (z-pass -> x-pass, zTOx) uses the same tiling with its own tile width twa (the
// lines one zfwd CTA transforms); twa < 0 selects the row-major layout [izl][ixl] of chb_buf_index.
CHB_HD size_t chb_bufA_index(int peer, int comp, int np, int pl, int nzB, int izl, int nxB, int ixl, int twa) {
    return twa < 0 ? chb_buf_index(peer, 3, comp, np, pl, nzB, izl, nxB, ixl)
                   : chb_bufB_index(peer, 3, comp, np, pl, nzB, izl, nxB, ixl, twa);
}

// mpi_transpose.f90:214-215 with npy=1
CHB_HD void chb_decompose(int nxp1, int nzd, int nranks, int rank, int* nx0, int* nxN, int* nz0, int* nzN) {
    *nx0 = rank * nxp1 / nranks;
    *nxN = (rank + 1) * nxp1 / nranks - 1;
    *nz0 = rank * nzd / nranks;
    *nzN = (rank + 1) * nzd / nranks - 1;
}
