// bodyforce_kernels.cu - the body-force hooks of the reference (body_forces/*/*.inc) as masked linear
// maps of the velocity, and the ghost-node extension of F at the start of buildrhs (dnsdata.f90:616-629).
#include "chb_internal.h"

// F_r = sum_c A[r][c] * V_c inside the mask (set_body_force of body_forces/*/*.inc): mask_y(iy)*mask_z(iz) for
// the coriolis hook (coriolis.inc:29-41), a general mask(iy,iz) for the am_f1 / am_butterfly hooks whose
// active region is not a product of a y- and a z-range (am_f1.inc:13-28, am_butterfly.inc:11-29)
__global__ void body_force_kernel(const cplx* __restrict__ V, cplx* __restrict__ F, Geometry g, BodyForce bf) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int iyp = blockIdx.y;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    if (bf.mask_yz) {
        if (bf.mask_yz[(size_t)iyp * g.nzt + izp] == 0.0) return;
    } else if (bf.mask_y[iyp] == 0.0 || bf.mask_z[izp] == 0.0) {
        return;
    }
    if (bf.exclude_mean && g.nx0 + ixl == 0 && izp == g.nz) return;
    const size_t plane = (size_t)g.M, comp = (size_t)g.nyp * plane;
    const size_t off = (size_t)iyp * plane + m;
    const cplx u = V[off], v = V[comp + off], w = V[2 * comp + off];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // a component the hook never assigns (an all-zero row of A: F3 of the coriolis force) keeps its value
        if (bf.A[r * 3 + 0] == 0.0 && bf.A[r * 3 + 1] == 0.0 && bf.A[r * 3 + 2] == 0.0) continue;
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

#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)   // the kernel-only emulation harnesses (tests/host_emul) stop here
void launch_body_force(chb_handle_s* h) {
    const Geometry& g = h->g;
    dim3 grid((unsigned)((g.M + 255) / 256), g.nyp);
    CHB_LAUNCH(grid, 256, 0, h->stream, body_force_kernel)(h->V, h->F, g, h->bf);
    h->launches++;
}
void launch_force_ghosts(chb_handle_s* h) {
    const Geometry& g = h->g;
    CHB_LAUNCH((unsigned)((g.M + 255) / 256), 256, 0, h->stream, force_ghost_kernel)(h->F, g, h->tab);
    h->launches++;
}
#endif
