"""Test infrastructure: run `-m gpu` tests against the fully emulated library (tests/host_emul/build_full_emul.sh:
the product sources compiled with g++, kernels on the CTA emulator, a host-memory CUDA runtime) instead of
libchannel_b200.so.  Usage: python tests/host_emul/run_gpu_tests_emulated.py <pytest args>.  Only this runner ever
points the ctypes loader at the emulated library; the package itself has no such switch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import channel_b200._lib as L  # noqa: E402

L.LIB_PATH = os.path.join(ROOT, "tests", "host_emul", "_build", "libchannel_b200_emul.so")
import pytest  # noqa: E402

sys.exit(pytest.main(["-m", "gpu", "-p", "no:cacheprovider"] + sys.argv[1:]))
