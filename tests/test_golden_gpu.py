"""CUDA path (through the C ABI) against the committed golden fixtures tests/golden/*.npz
(produced by the numpy oracle, tests/golden/make_golden.py).  The GPU box has no /root/reference
and needs none: fixtures + oracle travel with the repo."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from channel_b200 import Channel, DnsIn, RK1_rai
from tests.helpers import relerr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    return mg.CASES


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_gpu_reproduces_golden(path):
    name = os.path.splitext(os.path.basename(path))[0]
    pk, fk, cor, nsteps = _cases()[name]
    g = np.load(path)
    p = DnsIn(**pk)
    ch = Channel(p)                       # host tables from host_tables.cpp (not the oracle's)
    ch.upload_V(g["V0"])
    if cor:
        ch.config_coriolis(0.02, 9999999.0, 1.0)
    ch.cfl_prepass()
    lines = [ch.outstats()]
    # first substep intermediates on a second handle
    ch2 = Channel(p); ch2.capture_products(); ch2.upload_V(g["V0"])
    if cor:
        ch2.config_coriolis(0.02, 9999999.0, 1.0)
    ch2.cfl_prepass(); ch2.outstats()
    ch2.buildrhs(RK1_rai, False)
    P = ch2.download_products()
    for k in range(6):
        assert relerr(P[k], g["products"][k]) < 1e-12
    rhs = ch2.download_rhs()
    sl = slice(2, p.ny + 1)
    assert relerr(rhs[0, sl], g["rhs"][0, sl]) < 1e-12 and relerr(rhs[1, sl], g["rhs"][1, sl]) < 1e-12
    ch2.close()
    for i in range(nsteps):
        lines.append(ch.step())
        if i == 0:
            V1 = ch.download_V()
            for c in range(3):
                assert relerr(V1[c], g["V1"][c]) < 1e-12          # single step: 1e-12 relative
    Vend = ch.download_V()
    for c in range(3):
        assert relerr(Vend[c], g["Vend"][c]) < 1e-9               # <= 10 steps: 1e-9
    assert np.allclose(np.array(lines)[:, 1:9], g["lines"][:, 1:9], rtol=1e-8, atol=1e-10)   # Runtimedata cols 2-9
    assert np.allclose(np.array(lines)[:, [0, 9, 10]], g["lines"][:, [0, 9, 10]], rtol=1e-9)
    ch.close()
