/* channel_b200.h - C ABI of the B200-native hot path of davecats/channel.
 *
 * The reference has no FFI: the boundary is the Fortran module-procedure
 * interface between PROGRAM channel (channel.f90) and MODULE dnsdata
 * (dnsdata.f90), with state shared through module variables.  This header
 * declares the extern "C" entry points an iso_c_binding shim binds so that the
 * bodies of init_fft / convolutions / buildrhs / linsolve / vetaTOuvw /
 * computeflowrate run on the GPU while channel.f90 stays verbatim
 * (fortran/channel_b200_mod.f90).
 *
 * Conventions: plain pointers and sizes only; all reals are double; complex is
 * double[2] (re,im); host pointers unless the name ends in _d; every function
 * returns 0 on success or a non-zero error code (text from chb_last_error());
 * one handle per rank (= per GPU); calls on a handle are not thread-safe; on
 * nranks>1 every call is collective over the ranks, in the same order (the
 * reference's MPI convention).  Work is enqueued on the handle's stream; calls
 * that return host scalars synchronise.
 *
 * All file:line citations are into the reference tree.
 */
#ifndef CHANNEL_B200_H
#define CHANNEL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct chb_handle_s* chb_handle;

#define CHB_NCCL_ID_BYTES 128

/* Text of the last error on this thread. */
const char* chb_last_error(void);

/* Library/ABI version (major*1000+minor). */
int chb_version(void);

/* Fill `id` (CHB_NCCL_ID_BYTES) with an NCCL unique id; rank 0 calls it and the
 * host side broadcasts it (MPI_Bcast in the Fortran shim) before chb_create. */
int chb_get_nccl_unique_id(char* id);

/* Replaces init_MPI (x-decomposition only: npy must be 1; mpi_transpose.f90:175-272),
 * init_fft (ffts.f90:35-76) and the device part of init_memory (dnsdata.f90:129-177).
 * Computes nx0..nxN / nz0..nzN exactly as mpi_transpose.f90:214-215; requires
 * nranks | (nx+1) and nranks | nzd (README.md:154).  ni is 1/Re (dnsdata.f90:115).
 * nccl_id may be NULL when nranks==1.  device = CUDA device ordinal. */
int chb_create(chb_handle* h, int nx, int ny, int nz, int nxd, int nzd,
               double alfa0, double beta0, double ni, double a, double ymin, double ymax,
               int rank, int nranks, const char* nccl_id, int device);

/* Replaces free_fft (ffts.f90:110-115) and free_memory (dnsdata.f90:222-229). */
int chb_destroy(chb_handle h);

/* Local index ranges of this rank (mpi_transpose.f90:214-215). */
int chb_get_decomposition(chb_handle h, int* nx0, int* nxN, int* nz0, int* nzN);

/* Consumes the output of setup_derivatives (dnsdata.f90:241-286) and
 * setup_boundary_conditions (dnsdata.f90:290-308), which stay on the host, so the
 * coefficients are bit-identical to the caller's.
 *   y[ny+3]            y(-1:ny+1)
 *   d0,d1,d2,d4        der(1:ny-1)%d*(-2:2), row-major [(ny-1)][5]
 *   d140..d24np1       wall stencils (-2:2)
 *   v0bc..etanp1bc     BC vectors (-2:2)
 *   D0mat              D0mat(1:ny+1,-2:2) AFTER LU5decompStep, row-major [(ny+1)][5] */
int chb_set_tables(chb_handle h, const double* y,
                   const double* d0, const double* d1, const double* d2, const double* d4,
                   const double* d140, const double* d14m1, const double* d240, const double* d24m1,
                   const double* d14n, const double* d14np1, const double* d24n, const double* d24np1,
                   const double* v0bc, const double* v0m1bc, const double* vnbc, const double* vnp1bc,
                   const double* eta0bc, const double* eta0m1bc, const double* etanbc, const double* etanp1bc,
                   const double* D0mat);

/* Host <-> device field transfer, host layout = Fortran V(-1:ny+1,-nz:nz,nx0:nxN,1:3)
 * (complex; iy fastest), i.e. what read_restart_file fills (dnsdata.f90:677-720) and
 * save_restart_file writes (dnsdata.f90:821-848). */
int chb_upload_V(chb_handle h, const double* V_host);
int chb_download_V(chb_handle h, double* V_host);

/* Page-lock / unlock a caller-owned host array (the Fortran V) so the transfers above run at
 * full PCIe rate and asynchronously (cudaHostRegister / cudaHostUnregister). */
int chb_host_register(void* ptr, size_t bytes);
int chb_host_unregister(void* ptr);

/* Same transfer in the device layout [c][iy+1][ix-nx0][iz+nz] (no transposition). */
int chb_upload_V_planes(chb_handle h, const double* V_host);
int chb_download_V_planes(chb_handle h, double* V_host);

/* bc0(0,0)%u=u0; bcn(0,0)%u=uN  (channel.f90:122-124). */
int chb_set_wall_velocity(chb_handle h, double u0, double uN);

/* Module variables dnsdata.f90:28-32 that the path reads; meanpx is updated on the
 * device when CPI is on (linsolve_blocking.inc:87-97). */
int chb_set_forcing(chb_handle h, double meanpx, double meanpz, double meanflowx, double meanflowz,
                    int CPI, int CPI_type, double gamma);

/* channel.f90:96-115: CFL over planes 1..ny-1 (convolutions with compute_cfl),
 * flow rates fr(1:2) of the mean profile and the CPI update of meanpx. */
int chb_cfl_prepass(chb_handle h);

/* Body force, device path for the masked linear forces the reference ships
 * (body_forces/coriolis/coriolis.inc:29-41; the am hooks use the _yz variant below):
 *   F_r(iy,iz,ix) = sum_c A[r][c] * mask(iy,|iz| or iz) * V_c(iy,iz,ix)
 * A is 3x3 row-major; mask_y[ny+3] and mask_z[2nz+1] are 0/1 factors whose product
 * is the mask; exclude_mean!=0 leaves F(:,0,0,:) untouched.  Entries outside the
 * mask keep their previous value (the hooks only assign inside the mask).
 * enable=0 switches the body force off (bodyforce undefined, header.h:29). */
int chb_set_body_force_linear(chb_handle h, int enable, const double* A, const double* mask_y,
                              const double* mask_z, int exclude_mean);
/* Same with a general 0/1 (or weighting) mask over (iy, iz), row-major mask_yz[(ny+3)][(2nz+1)]: the am_f1
 * and am_butterfly hooks, whose active region couples y and z (am_f1.inc:13-28: lambda_z+ > 2.3 (y+)^2;
 * am_butterfly.inc:11-29: two boxes).  Replaces a separable mask set before, and vice versa. */
int chb_set_body_force_linear_yz(chb_handle h, int enable, const double* A, const double* mask_yz,
                                 int exclude_mean);
/* Generic path for arbitrary set_body_force hooks that are not masked linear maps: the caller evaluates its hook on
 * the host (chb_download_V, its own code) and uploads the result before chb_buildrhs; chb_set_body_force is then a
 * no-op.  Host layout = Fortran F(-1:ny+1,-nz:nz,nx0:nxN,1:3); chb_download_F returns the field in the same layout
 * (Force.cart.<n>.out, dnsdata.f90:905-906).  Slow by construction (two PCIe crossings per RK substep). */
int chb_upload_F(chb_handle h, const double* F_host);
int chb_download_F(chb_handle h, double* F_host);
/* set_body_force() call sites channel.f90:129-131,142-144,155-157. */
int chb_set_body_force(chb_handle h);

/* Replaces buildrhs (dnsdata.f90:611-673) including every convolutions call
 * (dnsdata.f90:487-602).  ode = RK?_rai(1:3).  Leaves the RHS of the eta- and
 * D2v-equations in V(1:ny-1,:,:,1:2), as the reference does (dnsdata.f90:667-671; V holds
 * velocities again after chb_linsolve), and accumulates cfl (max) when compute_cfl!=0. */
int chb_buildrhs(chb_handle h, const double* ode, double deltat, int compute_cfl);

/* Replaces linsolve (linsolve_blocking.inc:3-107, blocking semantics: includes the
 * mean-mode flow-rate / CPI update and the inline vetaTOuvw). lambda = RK(1)/deltat. */
int chb_linsolve(chb_handle h, double lambda);

/* nonblockingY split (channel.f90:137-139, linsolve_nonblocking.inc:75-159): no-ops,
 * chb_linsolve already did both; provided so either header.h variant links. */
int chb_vetaTOuvw(chb_handle h);
int chb_computeflowrate(chb_handle h, double lambda);

/* One full RK3 step (channel.f90:125-166) = 3 x (set_body_force, buildrhs, linsolve). */
int chb_rk3_step(chb_handle h, double deltat);

/* What outstats reads (dnsdata.f90:861-878); one small D2H + stream sync.
 * cfl is the max over ranks and is reset to 0 on the device (dnsdata.f90:861).
 * U_lo/W_lo = Re V(-1:3,0,0,{1,3}); U_hi/W_hi = Re V(ny-3:ny+1,0,0,{1,3}) (valid on
 * the rank with nx0==0, broadcast to all). */
int chb_get_step_scalars(chb_handle h, double* cfl, double* fr, double* corrpx, double* corrpz,
                         double* meanpx, double* meanpz,
                         double* U_lo, double* U_hi, double* W_lo, double* W_hi);

/* ---- restart / snapshot files ------------------------------------------
 * Replaces save_restart_file (dnsdata.f90:821-848; call sites dnsdata.f90:885,903,906 and
 * channel.f90:181) for the device-resident field: writes the reference's Dati.cart.out format
 * (68-byte header int32 nx,ny,nz + float64 alfa0,beta0,ni,a,ymin,ymax,time, then
 * V(-1:ny+1,-nz:nz,0:nx,1:3) in Fortran order) without a host copy of V.  Collective: every rank
 * pwrite()s its x-slab at the offset the reference's MPI-IO subarray view gives it
 * (mpi_transpose.f90:249-258), rank 0 (has_terminal) also writes the header; all ranks name the same
 * file.  field = 0 writes V, 1 the body force F (Force.cart.*.out, dnsdata.f90:906).
 * async_mode = 0: returns when this rank's bytes are written.  async_mode = 1: returns as soon as the
 * device holds a private copy of the field in file order (a few ms); the PCIe copy and the file
 * writes continue on a copy stream and a writer thread while the caller goes on time-stepping.  One
 * snapshot is in flight at a time (a second call first waits for the previous one).  Errors: 4 =
 * file I/O, 5 = no device memory for the asynchronous snapshot buffer. */
int chb_save_restart_file(chb_handle h, const char* filename, double time, int field, int async_mode);
/* Wait for an asynchronous snapshot; returns its error code (0 = written). */
int chb_restart_wait(chb_handle h);
/* Last snapshot of this rank: bytes written, device time of the layout transposition (ms), wall time from
 * the call to the last byte written (s).  Waits for an asynchronous snapshot first. */
int chb_restart_stats(chb_handle h, double* bytes, double* snapshot_ms, double* total_s);
/* Replaces read_restart_file (dnsdata.f90:677-704) when the file exists: checks the header against the
 * handle's nx,ny,nz,alfa0,beta0,ni,a,ymin,ymax (error 3 with the reference's message on mismatch,
 * dnsdata.f90:696-703), returns `time`, and loads this rank's x-slab into the device field (pread chunks
 * through pinned buffers overlapped with the PCIe copies, then the layout transposition on the device).
 * Error 4 if the file cannot be opened: the caller then generates its initial field (dnsdata.f90:705-719)
 * and uses chb_upload_V. */
int chb_read_restart_file(chb_handle h, const char* filename, double* time);

/* ---- convection-velocity diagnostic (the reference's #ifdef convvel) -------------------------
 * chb_set_convvel(h, 1) allocates Voldz and uconv (dnsdata.f90:141-143) and arms the diagnostic: the first chb_buildrhs
 * after every chb_get_step_scalars (= after every outstats, dnsdata.f90:858-860) then compares the z-transformed
 * velocities with those of the previous such sweep and accumulates cu = Im(conj(ust) dtu)/(ix alfa0 |ust|^2) into uconv
 * (dnsdata.f90:515-531,546-549).  chb_get_convvel returns uconv as [3][ny+3][nx+1][nzd] float64 (the order
 * save_convvel_file writes, dnsdata.f90:792-816) and convvel_cnt; chb_save_convvel_file does what outstats does at
 * the dt_field cadence (dnsdata.f90:908-913): uconv/convvel_cnt to the file, then uconv = 0, convvel_cnt = 0.
 * One GPU only in this version (error 2 otherwise). */
int chb_set_convvel(chb_handle h, int enable);
int chb_get_convvel(chb_handle h, double* uconv_host, long long* count);
int chb_save_convvel_file(chb_handle h, const char* filename);

/* ---- test / diagnostics accessors (not part of the Fortran binding) ---------- */
/* RHS left by chb_buildrhs: [2][ny+3][nxB][2nz+1] complex (0=eta, 1=D2v); rows 1..ny-1.  As in the reference
 * (dnsdata.f90:667-671) the RHS lives in V itself between chb_buildrhs and chb_linsolve: components 0 and 1 of rows
 * 1..ny-1 (the other rows of the returned array are the old velocities). */
int chb_download_rhs(chb_handle h, double* rhs_host);
/* The spectral products VVdz exist for one chunk of planes at a time (the reference keeps a 5-slot ring, ffts.f90:42).
 * chb_debug_capture_products(h, 1) makes every following sweep of chb_buildrhs keep a copy of all planes (6 complex per
 * point of extra device memory); chb_download_products then returns [6][ny+3][nxB][2nz+1] complex
 * = VVdz(izd(iz)+1, ix+1-nx0, 1:6, slot) for every plane (dnsdata.f90:609) of the last sweep. */
int chb_debug_capture_products(chb_handle h, int on);
int chb_download_products(chb_handle h, double* prod_host);
/* Body force field in the device layout [3][ny+3][nxB][2nz+1]. */
int chb_download_F_planes(chb_handle h, double* F_host);
/* Device timing of the kernels launched since the last reset: fills up to `cap`
 * entries of (name, total ms, launches); returns the number of distinct kernels. */
int chb_timing_enable(chb_handle h, int on);
int chb_timing_report(chb_handle h, char* names, int name_stride, double* ms, long long* launches, int cap);
/* Number of kernel launches issued by this handle since creation. */
long long chb_launch_count(chb_handle h);
/* Wait for everything enqueued on the handle's stream. */
int chb_sync(chb_handle h);
/* CUDA-event stopwatch on the handle's stream (the stream every kernel of this handle is
 * launched on): begin records an event, end records a second one, waits for it and returns
 * the elapsed device time in milliseconds. */
int chb_stopwatch_begin(chb_handle h);
int chb_stopwatch_end(chb_handle h, double* ms);
/* The handle's cudaStream_t (as void*), for callers that enqueue their own work behind it. */
int chb_get_stream(chb_handle h, void** stream);
/* Bytes of device memory this handle allocated. */
long long chb_device_bytes(chb_handle h);
/* Device ceilings measured with this library's own kernels: out[0] = FP64 FMA TFLOP/s,
 * out[1] = STREAM-style copy GB/s (read+write).  Diagnostics for the roofline report. */
int chb_measure_device_peaks(double* out);
/* Standalone batched FFT entry points used by the FFT parity tests:
 * complex lines of length n (sign=+1 backward / -1 forward, unnormalised), in place. */
int chb_test_fft_lines(int n, int nlines, int sign, double* data_host);

#ifdef __cplusplus
}
#endif
#endif
