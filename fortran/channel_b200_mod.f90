! channel_b200_mod.f90 - iso_c_binding shim: binds the C ABI of libchannel_b200.so
! (include/channel_b200.h) under the names PROGRAM channel already calls, so channel.f90 stays
! verbatim.  It replaces the BODIES of convolutions/buildrhs/linsolve/vetaTOuvw/computeflowrate
! and the device part of init_fft/init_memory in dnsdata.f90; read_dnsin, setup_derivatives,
! setup_boundary_conditions, read/save_restart_file, outstats and the body_forces hooks stay in
! dnsdata.f90 (see INTEGRATION.md for the exact patch).  Source only: this image has no Fortran
! compiler, so the shim is exercised through the equivalent ctypes host side
! (channel_b200/dnsdata.py), which makes the same calls in the same order.
MODULE channel_b200
  USE, INTRINSIC :: iso_c_binding
  IMPLICIT NONE
  TYPE(C_PTR), SAVE :: chb_h = C_NULL_PTR      ! one handle per MPI rank = per GPU

  INTERFACE
    FUNCTION chb_last_error() BIND(C, name="chb_last_error") RESULT(msg)
      IMPORT :: C_PTR
      TYPE(C_PTR) :: msg
    END FUNCTION
    FUNCTION chb_version() BIND(C, name="chb_version") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION chb_get_nccl_unique_id(id) BIND(C, name="chb_get_nccl_unique_id") RESULT(rc)
      IMPORT :: C_INT, C_CHAR
      CHARACTER(KIND=C_CHAR) :: id(128)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_create(h, nx, ny, nz, nxd, nzd, alfa0, beta0, ni, a, ymin, ymax, rank, nranks, nccl_id, device) &
        BIND(C, name="chb_create") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE, C_CHAR
      TYPE(C_PTR) :: h
      INTEGER(C_INT), VALUE :: nx, ny, nz, nxd, nzd, rank, nranks, device
      REAL(C_DOUBLE), VALUE :: alfa0, beta0, ni, a, ymin, ymax
      CHARACTER(KIND=C_CHAR) :: nccl_id(128)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_destroy(h) BIND(C, name="chb_destroy") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_get_decomposition(h, nx0, nxN, nz0, nzN) BIND(C, name="chb_get_decomposition") RESULT(rc)   ! mpi_transpose.f90:214-215
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT) :: nx0, nxN, nz0, nzN
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_set_tables(h, y, d0, d1, d2, d4, d140, d14m1, d240, d24m1, d14n, d14np1, d24n, d24np1, &
                            v0bc, v0m1bc, vnbc, vnp1bc, eta0bc, eta0m1bc, etanbc, etanp1bc, D0mat) &
        BIND(C, name="chb_set_tables") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE) :: y(*), d0(*), d1(*), d2(*), d4(*), d140(*), d14m1(*), d240(*), d24m1(*), d14n(*), d14np1(*), &
                        d24n(*), d24np1(*), v0bc(*), v0m1bc(*), vnbc(*), vnp1bc(*), eta0bc(*), eta0m1bc(*), etanbc(*), &
                        etanp1bc(*), D0mat(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_host_register(ptr, bytes) BIND(C, name="chb_host_register") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_SIZE_T
      TYPE(C_PTR), VALUE :: ptr
      INTEGER(C_SIZE_T), VALUE :: bytes
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_host_unregister(ptr) BIND(C, name="chb_host_unregister") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: ptr
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_upload_V(h, V) BIND(C, name="chb_upload_V") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE_COMPLEX
      TYPE(C_PTR), VALUE :: h
      COMPLEX(C_DOUBLE_COMPLEX) :: V(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_download_V(h, V) BIND(C, name="chb_download_V") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE_COMPLEX
      TYPE(C_PTR), VALUE :: h
      COMPLEX(C_DOUBLE_COMPLEX) :: V(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_set_wall_velocity(h, u0, uN) BIND(C, name="chb_set_wall_velocity") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE), VALUE :: u0, uN
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_set_forcing(h, meanpx, meanpz, meanflowx, meanflowz, CPI, CPI_type, gamma) &
        BIND(C, name="chb_set_forcing") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE), VALUE :: meanpx, meanpz, meanflowx, meanflowz, gamma
      INTEGER(C_INT), VALUE :: CPI, CPI_type
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_cfl_prepass(h) BIND(C, name="chb_cfl_prepass") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_set_body_force_linear(h, enable, A, mask_y, mask_z, exclude_mean) &
        BIND(C, name="chb_set_body_force_linear") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT), VALUE :: enable, exclude_mean
      REAL(C_DOUBLE) :: A(9), mask_y(*), mask_z(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_set_body_force_linear_yz(h, enable, A, mask_yz, exclude_mean) &
        BIND(C, name="chb_set_body_force_linear_yz") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT), VALUE :: enable, exclude_mean
      REAL(C_DOUBLE) :: A(9), mask_yz(*)      ! mask_yz(iz+nz+1 + (2nz+1)*(iy+1)): row-major (iy, iz)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_upload_F(h, F) BIND(C, name="chb_upload_F") RESULT(rc)   ! F(-1:ny+1,-nz:nz,nx0:nxN,1:3) from a host-side hook
      IMPORT :: C_PTR, C_INT, C_DOUBLE_COMPLEX
      TYPE(C_PTR), VALUE :: h
      COMPLEX(C_DOUBLE_COMPLEX) :: F(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_download_F(h, F) BIND(C, name="chb_download_F") RESULT(rc)   ! Force.cart.<n>.out through the driver's own writer
      IMPORT :: C_PTR, C_INT, C_DOUBLE_COMPLEX
      TYPE(C_PTR), VALUE :: h
      COMPLEX(C_DOUBLE_COMPLEX) :: F(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_set_body_force(h) BIND(C, name="chb_set_body_force") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_buildrhs(h, ode, deltat, compute_cfl) BIND(C, name="chb_buildrhs") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE) :: ode(3)
      REAL(C_DOUBLE), VALUE :: deltat
      INTEGER(C_INT), VALUE :: compute_cfl
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_linsolve(h, lambda) BIND(C, name="chb_linsolve") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE), VALUE :: lambda
      INTEGER(C_INT) :: rc
    END FUNCTION
    ! the nonblockingY variant of header.h calls vetaTOuvw and computeflowrate separately (channel.f90:137-139,
    ! linsolve_nonblocking.inc:75-159); chb_linsolve has done both, these return at once
    FUNCTION chb_vetaTOuvw(h) BIND(C, name="chb_vetaTOuvw") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_computeflowrate(h, lambda) BIND(C, name="chb_computeflowrate") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE), VALUE :: lambda
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_rk3_step(h, deltat) BIND(C, name="chb_rk3_step") RESULT(rc)   ! the three substeps of channel.f90:125-166 in one call
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE), VALUE :: deltat
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_get_step_scalars(h, cfl, fr, corrpx, corrpz, meanpx, meanpz, U_lo, U_hi, W_lo, W_hi) &
        BIND(C, name="chb_get_step_scalars") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE) :: cfl, fr(3), corrpx, corrpz, meanpx, meanpz, U_lo(5), U_hi(5), W_lo(5), W_hi(5)
      INTEGER(C_INT) :: rc
    END FUNCTION
    ! restart / snapshot files from the device-resident field (save_restart_file dnsdata.f90:821-848,
    ! read_restart_file dnsdata.f90:677-704); filename is a NUL-terminated C string: TRIM(filename)//C_NULL_CHAR
    FUNCTION chb_save_restart_file(h, filename, time, field, async_mode) BIND(C, name="chb_save_restart_file") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE, C_CHAR
      TYPE(C_PTR), VALUE :: h
      CHARACTER(KIND=C_CHAR) :: filename(*)
      REAL(C_DOUBLE), VALUE :: time
      INTEGER(C_INT), VALUE :: field, async_mode
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_restart_wait(h) BIND(C, name="chb_restart_wait") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_restart_stats(h, bytes, snapshot_ms, total_s) BIND(C, name="chb_restart_stats") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE) :: bytes, snapshot_ms, total_s
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_read_restart_file(h, filename, time) BIND(C, name="chb_read_restart_file") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE, C_CHAR
      TYPE(C_PTR), VALUE :: h
      CHARACTER(KIND=C_CHAR) :: filename(*)
      REAL(C_DOUBLE) :: time
      INTEGER(C_INT) :: rc
    END FUNCTION
    ! convection-velocity diagnostic (#ifdef convvel): arm once, save at the dt_field cadence of outstats
    FUNCTION chb_set_convvel(h, enable) BIND(C, name="chb_set_convvel") RESULT(rc)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT), VALUE :: enable
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_get_convvel(h, uconv, count) BIND(C, name="chb_get_convvel") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_DOUBLE, C_LONG_LONG
      TYPE(C_PTR), VALUE :: h
      REAL(C_DOUBLE) :: uconv(*)
      INTEGER(C_LONG_LONG) :: count
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION chb_save_convvel_file(h, filename) BIND(C, name="chb_save_convvel_file") RESULT(rc)
      IMPORT :: C_PTR, C_INT, C_CHAR
      TYPE(C_PTR), VALUE :: h
      CHARACTER(KIND=C_CHAR) :: filename(*)
      INTEGER(C_INT) :: rc
    END FUNCTION
  END INTERFACE

CONTAINS

  SUBROUTINE chb_check(rc, what)
    INTEGER(C_INT), INTENT(IN) :: rc
    CHARACTER(LEN=*), INTENT(IN) :: what
    IF (rc /= 0) THEN
      WRITE(*,*) "channel_b200: ", what, " failed with code ", rc   ! text: chb_last_error()
      STOP 1                                                        ! the reference STOPs on errors
    END IF
  END SUBROUTINE

END MODULE channel_b200
