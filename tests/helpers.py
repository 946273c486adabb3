"""Shared helpers for the parity tests: build an oracle and a GPU Channel on the same inputs."""
import numpy as np

from channel_b200 import Channel, DnsIn
from channel_b200.fields import perturbed_laminar
from oracle.channel_oracle import DnsIn as ODnsIn, Oracle


def relerr(a, b):
    """norm-wise relative error max|a-b| / max|b| (how utilities/compare_fields.py:17-54 compares fields)."""
    d = float(np.abs(a - b).max())
    n = float(np.abs(b).max())
    return d / n if n > 0 else d


def make_pair(nx, ny, nz, deltat=2e-3, cflmax=0.0, re=2000.0, eps=1e-2, seed=20261017, couette=False, **kw):
    fields = dict(nx=nx, ny=ny, nz=nz, re=re, deltat=deltat, cflmax=cflmax)
    fields.update(kw)
    p = DnsIn(**fields)
    op = ODnsIn(**{k: getattr(p, k) for k in ODnsIn.__dataclass_fields__})
    o = Oracle(op)
    V0 = perturbed_laminar(nx, ny, nz, p.alfa0, p.beta0, p.a, p.ymin, p.ymax, eps=eps, seed=seed, couette=couette)
    o.V[:] = V0
    ch = Channel(p, tables=o)        # identical coefficient tables on both sides
    ch.capture_products()            # the tests look at the spectral products of all planes (chb_download_products)
    ch.upload_V(V0)
    return p, o, ch, V0
