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
#include "chb_internal.h"
#include "solve_device.cuh"

#define RHS_THREADS 128

struct RhsAcc {
    cplx ev, ee, lv, le;
};

template <bool HAS_F>
__global__ void __launch_bounds__(RHS_THREADS)
rhs_kernel(const cplx* __restrict__ V, const cplx* __restrict__ P, const cplx* __restrict__ F, cplx* __restrict__ rhs,
           cplx* __restrict__ oldrhs, Geometry g, DevTables tab, const DevScalars* __restrict__ sc, double ode1_dt,
           double ode2, double ode3) {
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

    for (int ip = -1; ip <= ny + 1; ++ip) {
        const size_t off = (size_t)(ip + 1) * plane + m;
        const cplx p1 = P[0 * comp + off], p2 = P[1 * comp + off], p3 = P[2 * comp + off];
        const cplx p4 = P[3 * comp + off], p5 = P[4 * comp + off], p6 = P[5 * comp + off];
        const cplx u = V[0 * comp + off], v = V[1 * comp + off], w = V[2 * comp + off];
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
            const cplx oe = oldrhs[0 * comp + oo], ov = oldrhs[1 * comp + oo];
            cplx re, rv;
            re.x = acc[0].le.x + ode2 * ee.x - ode3 * oe.x;
            re.y = acc[0].le.y + ode2 * ee.y - ode3 * oe.y;
            rv.x = acc[0].lv.x + ode2 * acc[0].ev.x - ode3 * ov.x;
            rv.y = acc[0].lv.y + ode2 * acc[0].ev.y - ode3 * ov.y;
            rhs[0 * comp + oo] = re;
            rhs[1 * comp + oo] = rv;
            oldrhs[0 * comp + oo] = ee;
            oldrhs[1 * comp + oo] = acc[0].ev;
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[s] = acc[s + 1];
        acc[4].ev = acc[4].ee = acc[4].lv = acc[4].le = make_double2(0.0, 0.0);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused buildrhs plane loop + first sweep of linsolve (CHB_FUSE=1): the same scatter-form assembly,
// marching DOWN through the planes, so that every completed right-hand side (rows ny-1 .. 1) goes
// straight into the descending sweep of the banded UL solve (S1 of solve_kernels.cu: matrix rows of
// linsolve_blocking.inc:12-13, BC folding, LU5decompStep and LeftLU5divStep1 on the fly) without a
// round trip through HBM.  Writes the Step1 intermediates of both equations where solve_s2 expects
// them (`xout` = the rhs array) and the new explicit terms to oldrhs: 11 C read + 4.5 C written per
// point instead of (11 + 4) + (2 + 2.5).  lam = ODE(1)/deltat is the argument linsolve receives
// (channel.f90:130-138 pass the same RK coefficient to both).
template <bool HAS_F, int MINB>
__global__ void __launch_bounds__(RHS_THREADS, MINB)
rhs_s1_kernel(const cplx* __restrict__ V, const cplx* __restrict__ P, const cplx* __restrict__ F,
              cplx* __restrict__ xout, cplx* __restrict__ oldrhs, double* __restrict__ ckpt, Geometry g, DevTables tab,
              const DevScalars* __restrict__ sc, double lam, double ode2, double ode3) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.M) return;
    const int ixl = (int)(m / g.nzt);
    const int izp = (int)(m - (long long)ixl * g.nzt);
    const int ix = g.nx0 + ixl;
    const int iz = izp - g.nz;
    const double al = g.alfa0 * ix, be = g.beta0 * iz;
    const double k2 = al * al + be * be;
    const bool mean = (ix == 0 && iz == 0);
    const double ni = g.ni;
    const double cv0 = ni * k2 * k2 - lam * k2, cv2 = lam - 2.0 * ni * k2;
    const double ce0 = lam - ni * k2;
    const size_t plane = (size_t)g.M;
    const size_t comp = (size_t)g.nyp * plane;
    const int ny = g.ny;

    RhsAcc acc[5];   // acc[s] <-> output plane io = ip + 2 - s, stencil offset j = s - 2
#pragma unroll
    for (int s = 0; s < 5; ++s) acc[s].ev = acc[s].ee = acc[s].lv = acc[s].le = make_double2(0.0, 0.0);
    double mpx = 0.0, mpz = 0.0, bc0_eta = 0.0, bcn_eta = 0.0;
    if (mean) {
        mpx = sc->meanpx;
        mpz = sc->meanpz;
        bc0_eta = sc->u0;   // linsolve_blocking.inc:17,31: only the mean mode carries non-zero wall data
        bcn_eta = sc->uN;
    }
    LUState ste = {0, 0, 0, 0}, stv = {0, 0, 0, 0};
    cplx xe1 = make_double2(0, 0), xe2 = xe1, xv1 = xe1, xv2 = xe1;   // x(i+1), x(i+2) of both equations

    for (int ip = ny + 1; ip >= -1; --ip) {
        const size_t off = (size_t)(ip + 1) * plane + m;
        const cplx p1 = P[0 * comp + off], p2 = P[1 * comp + off], p3 = P[2 * comp + off];
        const cplx p4 = P[3 * comp + off], p5 = P[4 * comp + off], p6 = P[5 * comp + off];
        const cplx u = V[0 * comp + off], v = V[1 * comp + off], w = V[2 * comp + off];
        cplx f1 = make_double2(0, 0), f2 = f1, f3 = f1;
        if (HAS_F) {
            f1 = F[0 * comp + off];
            f2 = F[1 * comp + off];
            f3 = F[2 * comp + off];
        }
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
            gg.x = -be * u.y + al * w.y;
            gg.y = be * u.x - al * w.x;
        } else {
            T0e = make_double2(0.0, 0.0);
            T1e = make_double2(-p4.x, -p5.x);
            gg = make_double2(u.x, w.x);
        }
        if (HAS_F) {
            T0v.x -= k2 * f2.x;
            T0v.y -= k2 * f2.y;
            T1v.x += al * f1.y + be * f3.y;
            T1v.y -= al * f1.x + be * f3.x;
            if (!mean) {
                T0e.x += -be * f1.y + al * f3.y;
                T0e.y += be * f1.x - al * f3.x;
            } else {
                T0e.x += f1.x;
                T0e.y += f3.x;
            }
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int io = ip + 2 - s;
            if (io >= 1 && io <= ny - 1) {
                const int ti = (io + 1) * 5 + s;
                const double c0 = __ldg(&tab.d0[ti]), c1 = __ldg(&tab.d1[ti]);
                const double c2 = __ldg(&tab.d2[ti]), c4 = __ldg(&tab.d4[ti]);
                acc[s].ev.x += c0 * T0v.x + c1 * T1v.x + c2 * T2v.x;
                acc[s].ev.y += c0 * T0v.y + c1 * T1v.y + c2 * T2v.y;
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
        // output plane io = ip+2 is complete: timescheme (dnsdata.f90:486), then its row of the UL sweep
        const int io = ip + 2;
        if (io >= 1 && io <= ny - 1) {
            const size_t oo = (size_t)(io + 1) * plane + m;
            cplx ee = acc[0].ee;
            if (mean) {
                ee.x += mpx;
                ee.y += mpz;
            }
            const cplx oe = oldrhs[0 * comp + oo], ov = oldrhs[1 * comp + oo];
            cplx be_, bv_;   // right-hand sides of the eta and v equations
            be_.x = acc[0].le.x + ode2 * ee.x - ode3 * oe.x;
            be_.y = acc[0].le.y + ode2 * ee.y - ode3 * oe.y;
            bv_.x = acc[0].lv.x + ode2 * acc[0].ev.x - ode3 * ov.x;
            bv_.y = acc[0].lv.y + ode2 * acc[0].ev.y - ode3 * ov.y;
            oldrhs[0 * comp + oo] = ee;
            oldrhs[1 * comp + oo] = acc[0].ev;
            Row5 rv, re;
            build_rows(tab, io, k2, lam, ni, rv, re);
            if (io == ny - 1) {
                fold_top1(rv, tab.vnbc, tab.vnp1bc);
                fold_top1(re, tab.etanbc, tab.etanp1bc);
                be_.x -= re.a[3] * bcn_eta / tab.etanbc[3];   // linsolve_blocking.inc:40
                rv.a[3] = rv.a[4] = 0.0;                      // rbparmat_blocking.f90:29
                re.a[3] = re.a[4] = 0.0;
            } else if (io == ny - 2) {
                fold_top2(rv, tab.vnbc);
                fold_top2(re, tab.etanbc);
                be_.x -= re.a[4] * bcn_eta / tab.etanbc[3];   // :41
                rv.a[4] = 0.0;
                re.a[4] = 0.0;
            }
            if (io == 1) {
                fold_bot1(rv, tab.v0bc, tab.v0m1bc);
                fold_bot1(re, tab.eta0bc, tab.eta0m1bc);
                be_.x -= re.a[1] * bc0_eta / tab.eta0bc[1];   // :26
            } else if (io == 2) {
                fold_bot2(rv, tab.v0bc);
                fold_bot2(re, tab.eta0bc);
                be_.x -= re.a[0] * bc0_eta / tab.eta0bc[1];   // :27
            }
            double inv, u1, u2;
            lu_row(re, ste, inv, u1, u2);
            cplx x;                                           // LeftLU5divStep1 (rbparmat_blocking.f90:70-72)
            x.x = (be_.x - (u1 * xe1.x + u2 * xe2.x)) * inv;
            x.y = (be_.y - (u1 * xe1.y + u2 * xe2.y)) * inv;
            xe2 = xe1;
            xe1 = x;
            xout[0 * comp + oo] = x;
            lu_row(rv, stv, inv, u1, u2);
            x.x = (bv_.x - (u1 * xv1.x + u2 * xv2.x)) * inv;
            x.y = (bv_.y - (u1 * xv1.y + u2 * xv2.y)) * inv;
            xv2 = xv1;
            xv1 = x;
            xout[1 * comp + oo] = x;
            if (io > 1 && (io - 1) % SOLVE_K == 0) {          // checkpoints of the UL recurrence for solve_s2
                double* ck = ckpt + (size_t)((io - 1) / SOLVE_K - 1) * 8 * plane + m;
                ck[0 * plane] = stv.l1m2; ck[1 * plane] = stv.l1m1; ck[2 * plane] = stv.l2m2; ck[3 * plane] = stv.l2m1;
                ck[4 * plane] = ste.l1m2; ck[5 * plane] = ste.l1m1; ck[6 * plane] = ste.l2m2; ck[7 * plane] = ste.l2m1;
            }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[s] = acc[s + 1];
        acc[4].ev = acc[4].ee = acc[4].lv = acc[4].le = make_double2(0.0, 0.0);
    }
}

#ifndef CHB_HOST_EMUL   // tests/host_emul compiles the kernels above with g++ and runs them thread by thread
void launch_rhs_s1(chb_handle_s* h, const double* ode, double deltat) {
    const Geometry& g = h->g;
    const int blocks = (int)((g.M + RHS_THREADS - 1) / RHS_THREADS);
    ScopedKernelTimer tm(h, "rhs_s1");
    // CHB_FUSE=1: 2 blocks/SM, no spills; CHB_FUSE=2: 168-register cap, 3 blocks/SM, 80 bytes of spills
    auto kern = h->bf.enabled ? (h->fuse == 2 ? rhs_s1_kernel<true, 3> : rhs_s1_kernel<true, 2>)
                              : (h->fuse == 2 ? rhs_s1_kernel<false, 3> : rhs_s1_kernel<false, 2>);
    kern<<<blocks, RHS_THREADS, 0, h->stream>>>(h->V, h->P, h->bf.enabled ? h->F : nullptr, h->rhs, h->oldrhs, h->ckpt, g,
                                               h->tab, h->sc, ode[0] / deltat, ode[1], ode[2]);
    h->launches++;
}

void launch_rhs(chb_handle_s* h, const double* ode, double deltat) {
    const Geometry& g = h->g;
    const int blocks = (int)((g.M + RHS_THREADS - 1) / RHS_THREADS);
    ScopedKernelTimer tm(h, "rhs");
    if (h->bf.enabled)
        rhs_kernel<true><<<blocks, RHS_THREADS, 0, h->stream>>>(h->V, h->P, h->F, h->rhs, h->oldrhs, g, h->tab, h->sc,
                                                                ode[0] / deltat, ode[1], ode[2]);
    else
        rhs_kernel<false><<<blocks, RHS_THREADS, 0, h->stream>>>(h->V, h->P, nullptr, h->rhs, h->oldrhs, g, h->tab,
                                                                 h->sc, ode[0] / deltat, ode[1], ode[2]);
    h->launches++;
}
#endif
