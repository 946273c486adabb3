// rhs_kernel.cu - RHS assembly of the v- and eta-equations (buildrhs, dnsdata.f90:611-673).
//
// One thread per wavenumber column (ix,iz), marching through the y-planes once.  The 5-point
// y-stencils DD(f,k) of the reference (dnsdata.f90:609) are applied in scatter form: every input
// plane contributes to the five output planes iy-2..iy+2 that are still "in flight", so each
// product / velocity plane is read from HBM exactly once and the window lives in registers.
// Because ialfa, ibeta and k2 are per-column constants the explicit terms reduce to
//   expl_v   = sum_j d0(j) T0v + d1(j) T1v + d2(j) T2v      (dnsdata.f90:638-646)
//   expl_eta = sum_j d0(j) T0e + d1(j) T1e                  (dnsdata.f90:656-660)
// with per-plane combinations
//   T2v = ia*P4 + ib*P5                  T0v = k2*T2v            [- k2*F2]
//   T1v = ia*ia*P1 + 2*ia*ib*P6 + ib*ib*P3 + k2*P2               [- ia*F1 - ib*F3]
//   T0e = alfa*beta*(P1-P3) + (beta^2-alfa^2)*P6                 [+ ib*F1 - ia*F3]
//   T1e = -ib*P4 + ia*P5
// and the implicit / time-derivative parts (timescheme, dnsdata.f90:486; OS,SQ :476-477)
//   lin_v   = sum_j [ODE1/dt*(d2-k2*d0) + ni*(d4-2*k2*d2+k2^2*d0)](j) * v(iy+j)
//   lin_eta = sum_j [ODE1/dt*d0 + ni*(d2-k2*d0)](j) * (ib*u - ia*w)(iy+j)
// The mean mode (ix=iz=0) uses the real/imag packing of dnsdata.f90:648-654.
#include <cstdlib>
#include "chb_internal.h"
#include "solve_device.cuh"

#define RHS_THREADS 128

struct RhsAcc {
    cplx ev, ee, lv, le;
};

// MINB = resident blocks per SM the register allocation is sized for: 3 = 152 registers, no spills (4 = 128 registers
// with 56-96 bytes of spills measured no faster).
// The march covers the input planes ip0..ip1 of one chunk of the convolutions and carries the four partially
// accumulated output planes to the next launch through `state` ([32][M] doubles): the plane loop of buildrhs follows
// the convolutions chunk by chunk (and runs beside the transposes of the next chunk) instead of waiting for all
// planes, so the spectral products P exist for one chunk only ([6][pnp][M], plane index relative to pplane0).  Same
// operations in the same order as a single march.  The finished rows go where the reference puts them
// (dnsdata.f90:667-671): into V itself, eta-RHS -> component 0, D2v-RHS -> component 1 of row iy-2, two planes behind
// the march, where the old velocities are no longer needed (by this thread: same column; by the z-passes: earlier chunk).
template <bool HAS_F, int MINB, bool CG = false>
__global__ void __launch_bounds__(RHS_THREADS, MINB)
rhs_kernel(cplx* V, const cplx* __restrict__ P, const cplx* __restrict__ F,
           cplx* __restrict__ oldrhs, Geometry g, DevTables tab, const DevScalars* __restrict__ sc, double ode1_dt,
           double ode2, double ode3, int ip0, int ip1, double* __restrict__ state, int pnp, int pplane0) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    const int ix = g.nx0 + ixl;
    const int iz = izp - g.nz;
    const double al = g.alfa0 * ix, be = g.beta0 * iz;  // ialfa = i*al, ibeta = i*be (dnsdata.f90:156-157)
    const double k2 = al * al + be * be;                  // dnsdata.f90:158
    const bool mean = (ix == 0 && iz == 0);
    const double ni = g.ni;
    const double cv0 = ni * k2 * k2 - ode1_dt * k2, cv2 = ode1_dt - 2.0 * ni * k2;  // coefficient of d0, d2 in lin_v
    const double ce0 = ode1_dt - ni * k2;                                           // coefficient of d0 in lin_eta
    const size_t plane = (size_t)g.M;
    const size_t comp = (size_t)g.nyp * plane;
    const size_t pcomp = (size_t)pnp * plane;
    const int ny = g.ny;

    RhsAcc acc[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        acc[s].ev = acc[s].ee = acc[s].lv = acc[s].le = make_double2(0.0, 0.0);
    }
    double mpx = 0.0, mpz = 0.0;
    if (mean) {
        mpx = sc->meanpx;
        mpz = sc->meanpz;
    }
    {
        if (ip0 > -1) {   // resume: the four output planes ip0-2 .. ip0+1 are partially accumulated
            const double* st = state + m;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                acc[s].ev = make_double2(st[(8 * s + 0) * plane], st[(8 * s + 1) * plane]);
                acc[s].ee = make_double2(st[(8 * s + 2) * plane], st[(8 * s + 3) * plane]);
                acc[s].lv = make_double2(st[(8 * s + 4) * plane], st[(8 * s + 5) * plane]);
                acc[s].le = make_double2(st[(8 * s + 6) * plane], st[(8 * s + 7) * plane]);
            }
        }
    }
    for (int ip = ip0; ip <= ip1; ++ip) {
        const size_t off = (size_t)(ip + 1) * plane + m;
        const size_t poff = (size_t)(ip + 1 - pplane0) * plane + m;
        const cplx p1 = P[0 * pcomp + poff], p2 = P[1 * pcomp + poff], p3 = P[2 * pcomp + poff];
        const cplx p4 = P[3 * pcomp + poff], p5 = P[4 * pcomp + poff], p6 = P[5 * pcomp + poff];
        // read-only path for V: a plane is read two iterations before this thread overwrites it, nobody else writes it
        cplx u, v, w;
        if constexpr (CG) {   // L2 only: the rows are overwritten two planes later
            u = __ldcg(reinterpret_cast<const double2*>(V + 0 * comp + off));
            v = __ldcg(reinterpret_cast<const double2*>(V + 1 * comp + off));
            w = __ldcg(reinterpret_cast<const double2*>(V + 2 * comp + off));
        } else {
            u = __ldg(V + 0 * comp + off); v = __ldg(V + 1 * comp + off); w = __ldg(V + 2 * comp + off);
        }
        cplx f1 = make_double2(0, 0), f2 = f1, f3 = f1;
        if (HAS_F) {
            f1 = F[0 * comp + off];
            f2 = F[1 * comp + off];
            f3 = F[2 * comp + off];
        }
        // per-plane combinations (i*al*z = (-al*z.y, al*z.x))
        cplx T2v, T1v, T0v, T0e, T1e, gg;
        T2v.x = -(al * p4.y + be * p5.y);
        T2v.y = al * p4.x + be * p5.x;
        T1v.x = -(al * al) * p1.x - 2.0 * al * be * p6.x - (be * be) * p3.x + k2 * p2.x;
        T1v.y = -(al * al) * p1.y - 2.0 * al * be * p6.y - (be * be) * p3.y + k2 * p2.y;
        T0v.x = k2 * T2v.x;
        T0v.y = k2 * T2v.y;
        if (!mean) {
            T0e.x = al * be * (p1.x - p3.x) + (be * be - al * al) * p6.x;
            T0e.y = al * be * (p1.y - p3.y) + (be * be - al * al) * p6.y;
            T1e.x = be * p4.y - al * p5.y;
            T1e.y = -be * p4.x + al * p5.x;
            gg.x = -be * u.y + al * w.y;  // ib*u - ia*w
            gg.y = be * u.x - al * w.x;
        } else {
            // expl = (Re rhsu + meanpx) + i (Re rhsw + meanpz), rhsu = -DD(d1,4), rhsw = -DD(d1,5)
            T0e = make_double2(0.0, 0.0);
            T1e = make_double2(-p4.x, -p5.x);
            gg = make_double2(u.x, w.x);  // rD0(V,1,3): Re u + i Re w
        }
        if (HAS_F) {
            T0v.x -= k2 * f2.x;
            T0v.y -= k2 * f2.y;
            T1v.x += al * f1.y + be * f3.y;  // -ia*F1 - ib*F3
            T1v.y -= al * f1.x + be * f3.x;
            if (!mean) {
                T0e.x += -be * f1.y + al * f3.y;  // ib*F1 - ia*F3
                T0e.y += be * f1.x - al * f3.x;
            } else {
                T0e.x += f1.x;  // rD0(F,1,3)
                T0e.y += f3.x;
            }
        }
        // scatter to output planes io = ip-2+s, stencil offset j = ip-io = 2-s
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int io = ip - 2 + s;
            if (io >= 1 && io <= ny - 1) {
                const int ti = (io + 1) * 5 + (4 - s);
                const double c0 = __ldg(&tab.d0[ti]), c1 = __ldg(&tab.d1[ti]);
                const double c2 = __ldg(&tab.d2[ti]), c4 = __ldg(&tab.d4[ti]);
                const double c02 = c2;
                acc[s].ev.x += c0 * T0v.x + c1 * T1v.x + c02 * T2v.x;
                acc[s].ev.y += c0 * T0v.y + c1 * T1v.y + c02 * T2v.y;
                acc[s].ee.x += c0 * T0e.x + c1 * T1e.x;
                acc[s].ee.y += c0 * T0e.y + c1 * T1e.y;
                const double cv = cv0 * c0 + cv2 * c2 + ni * c4;
                const double ce = ce0 * c0 + ni * c2;
                acc[s].lv.x += cv * v.x;
                acc[s].lv.y += cv * v.y;
                acc[s].le.x += ce * gg.x;
                acc[s].le.y += ce * gg.y;
            }
        }
        // output plane io = ip-2 is complete: timescheme (dnsdata.f90:486)
        const int io = ip - 2;
        if (io >= 1 && io <= ny - 1) {
            const size_t oo = (size_t)(io + 1) * plane + m;
            cplx ee = acc[0].ee;
            if (mean) {
                ee.x += mpx;
                ee.y += mpz;
            }
            // RK substep 1 has ODE(3) = 0 (RK1_rai, dnsdata.f90:70): the previous explicit term is not read
            cplx oe = make_double2(0.0, 0.0), ov = oe;
            if (ode3 != 0.0) {
                oe = oldrhs[0 * comp + oo];
                ov = oldrhs[1 * comp + oo];
            }
            cplx re, rv;
            re.x = acc[0].le.x + ode2 * ee.x - ode3 * oe.x;
            re.y = acc[0].le.y + ode2 * ee.y - ode3 * oe.y;
            rv.x = acc[0].lv.x + ode2 * acc[0].ev.x - ode3 * ov.x;
            rv.y = acc[0].lv.y + ode2 * acc[0].ev.y - ode3 * ov.y;
            if constexpr (CG) {
                __stcg(reinterpret_cast<double2*>(V + 0 * comp + oo), re);
                __stcg(reinterpret_cast<double2*>(V + 1 * comp + oo), rv);
            } else {
                V[0 * comp + oo] = re;
                V[1 * comp + oo] = rv;
            }
            oldrhs[0 * comp + oo] = ee;
            oldrhs[1 * comp + oo] = acc[0].ev;
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[s] = acc[s + 1];
        acc[4].ev = acc[4].ee = acc[4].lv = acc[4].le = make_double2(0.0, 0.0);
    }
    {
        if (ip1 < ny + 1) {
            double* st = state + m;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                st[(8 * s + 0) * plane] = acc[s].ev.x; st[(8 * s + 1) * plane] = acc[s].ev.y;
                st[(8 * s + 2) * plane] = acc[s].ee.x; st[(8 * s + 3) * plane] = acc[s].ee.y;
                st[(8 * s + 4) * plane] = acc[s].lv.x; st[(8 * s + 5) * plane] = acc[s].lv.y;
                st[(8 * s + 6) * plane] = acc[s].le.x; st[(8 * s + 7) * plane] = acc[s].le.y;
            }
        }
    }
}


#if !defined(CHB_HOST_EMUL) || defined(CHB_HOST_EMUL_FULL)   // the kernel-only emulation harnesses (tests/host_emul) stop here
// one chunk of input planes [plane0, plane0 + nplanes) (plane index = iy + 1) whose products are in the lane in use, on
// stream `st`; chunks must be launched in ascending order on the same stream (the carried accumulators)
void launch_rhs_chunk(chb_handle_s* h, const double* ode, double deltat, int plane0, int nplanes, cudaStream_t st) {
    const Geometry& g = h->g;
    const int blocks = (int)((g.M + RHS_THREADS - 1) / RHS_THREADS);
    ScopedKernelTimer tm(h, "rhs", st);
    static const bool cg = []() { const char* e = getenv("CHB_RHS_CG"); return e ? atoi(e) != 0 : true; }();   // V rows through L2 only (see solve_kernels.cu)
    auto kern = cg ? (h->bf.enabled ? rhs_kernel<true, 3, true> : rhs_kernel<false, 3, true>)
                   : (h->bf.enabled ? rhs_kernel<true, 3> : rhs_kernel<false, 3>);
    CHB_LAUNCH(blocks, RHS_THREADS, 0, st, kern)(h->V, h->Pc, h->bf.enabled ? h->F : nullptr, h->oldrhs, g, h->tab, h->sc,
                                         ode[0] / deltat, ode[1], ode[2], plane0 - 1, plane0 + nplanes - 2, h->rhs_state,
                                         h->chunk_planes, plane0);
    h->launches++;
}
#endif
