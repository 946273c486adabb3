/* channel_b200_host.h - host-side helpers of the C++/Python restatement of the reference's
 * driver (PROGRAM channel) and of the parts of MODULE dnsdata that stay on the CPU.  A Fortran
 * caller does not need these: it computes the same tables with its own setup_derivatives /
 * setup_boundary_conditions and passes them to chb_set_tables.
 */
#ifndef CHANNEL_B200_HOST_H
#define CHANNEL_B200_HOST_H
#include "channel_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Caller-allocated tables (sizes as in chb_set_tables). */
typedef struct chb_host_tables {
    double* y;      /* [ny+3] */
    double* d0;     /* [(ny-1)*5] */
    double* d1;
    double* d2;
    double* d4;
    double* D0mat;  /* [(ny+1)*5] */
    double d140[5], d14m1[5], d240[5], d24m1[5], d14n[5], d14np1[5], d24n[5], d24np1[5];
    double v0bc[5], v0m1bc[5], vnbc[5], vnp1bc[5], eta0bc[5], eta0m1bc[5], etanbc[5], etanp1bc[5];
} chb_host_tables;

/* fftFIT (ffts.f90:78-86) and the nxd/nzd rule of read_dnsin (dnsdata.f90:110-113). */
int chb_host_fft_fit(int n);
int chb_host_padded_sizes(int nx, int nz, int* nxd, int* nzd);

/* Grid (dnsdata.f90:153), setup_derivatives (dnsdata.f90:241-286) and
 * setup_boundary_conditions (dnsdata.f90:290-308) for the full channel, npy=1. */
int chb_host_setup_tables(int ny, double a, double ymin, double ymax, chb_host_tables* t);

/* init_MPI's x/z index ranges for npy=1 (mpi_transpose.f90:214-215); no GPU needed. */
int chb_host_decomposition(int nx, int nzd, int nranks, int rank, int* nx0, int* nxN, int* nz0, int* nzN);

/* Element offset (in complex numbers) inside the pencil-transpose work buffers
 * buf[peer][comp][plane][izl][ixl] that the z-pass / x-pass kernels write and read and that
 * the all-to-all exchanges block-wise (zTOx / xTOz, mpi_transpose.f90:50-117). */
long long chb_host_transpose_index(int peer, int ncomp, int comp, int nplanes, int plane, int nzB, int izl,
                                   int nxB, int ixl);

/* Dati.cart.out layout (dnsdata.f90:683-695,830-846; mpi_transpose.f90:249-258), no GPU needed:
 * the 68-byte header, the byte offset of component c (0..2) of the x-slab that starts at mode nx0, and
 * the size of the whole file. */
int chb_host_restart_header(int nx, int ny, int nz, double alfa0, double beta0, double ni, double a,
                            double ymin, double ymax, double time, unsigned char* out68);
long long chb_host_restart_offset(int nx, int ny, int nz, int nx0, int c);
long long chb_host_restart_file_bytes(int nx, int ny, int nz);

/* chb_set_tables with the contents of *t. */
int chb_host_apply_tables(chb_handle h, const chb_host_tables* t);

#ifdef __cplusplus
}
#endif
#endif
