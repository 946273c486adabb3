"""Generates tests/golden/*.npz from the numpy oracle (oracle/channel_oracle.py).

The reference ships no golden vectors for this path and cannot be run in this image
(SURVEY.md 8c), so these fixtures pin the *oracle*: any later change to the oracle, to the
C restatement or to the CUDA path that alters the numbers shows up as a diff against them.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.channel_oracle import DnsIn, Oracle, coriolis_force, RK1_rai  # noqa: E402
from channel_b200.fields import perturbed_laminar  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (DnsIn kwargs, field kwargs, coriolis?, steps)
    "poiseuille_7x16x5": (dict(nx=7, ny=16, nz=5, re=2000.0, deltat=0.0, cflmax=1.0), dict(eps=1e-2), False, 10),
    "poiseuille_12x24x9_fixeddt": (dict(nx=12, ny=24, nz=9, re=3000.0, deltat=2e-3, cflmax=0.0), dict(eps=1e-2), False, 3),
    "couette_coriolis_9x20x6": (dict(nx=9, ny=20, nz=6, re=1500.0, deltat=2e-3, cflmax=0.0, CPI=False, u0=-1.0, uN=1.0),
                                  dict(eps=1e-2, couette=True), True, 3),
}


def run(name):
    pk, fk, cor, nsteps = CASES[name]
    p = DnsIn(**pk)
    o = Oracle(p)
    V0 = perturbed_laminar(p.nx, p.ny, p.nz, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, **fk)
    o.V[:] = V0
    if cor:
        o.set_body_force(coriolis_force(0.02, 9999999.0, 1.0))
    o.cfl_prepass()
    lines = [o.outstats()]
    # single-substep intermediates of the first substep: products and RHS
    o2 = Oracle(p); o2.V[:] = V0
    if cor:
        coriolis_force(0.02, 9999999.0, 1.0)(o2)
    o2.cfl_prepass(); o2.outstats()
    P = o2.convolutions(o2.V, False)[..., o2.izd]
    rhs = o2.buildrhs(RK1_rai, False)
    V1 = None
    for i in range(nsteps):
        lines.append(o.step())
        if i == 0:
            V1 = o.V.copy()
    out = dict(V0=V0, V1=V1, Vend=o.V.copy(), lines=np.array(lines), products=P, rhs=rhs,
               d0=o.d0, d1=o.d1, d2=o.d2, d4=o.d4, D0mat=o.D0mat, y=o.y, v0bc=o.v0bc, eta0m1bc=o.eta0m1bc,
               vnbc=o.vnbc, etanp1bc=o.etanp1bc)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written;", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    for n in CASES:
        run(n)
