"""FFT kernels (replacement of ffts.f90 / FFTW3) against numpy.fft, FP64, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("sign", [1, -1, 2, -2])
def test_fft_lines_match_numpy(lib, n, sign):
    from channel_b200 import _lib
    rng = np.random.default_rng(n * 7 + sign)
    nlines = 13
    x = rng.standard_normal((nlines, n)) + 1j * rng.standard_normal((nlines, n))
    y = np.ascontiguousarray(x.copy())
    _lib.check(lib.chb_test_fft_lines(n, nlines, sign, y.ctypes.data), "chb_test_fft_lines")
    ref = np.fft.ifft(x, axis=1, norm="forward") if sign > 0 else np.fft.fft(x, axis=1)
    err = np.abs(y - ref).max() / np.abs(ref).max()
    assert err < 1e-14 * max(4, np.log2(n)), (n, sign, err)   # tolerance: a few ulp per pass


def test_fft_rejects_bad_size(lib):
    x = np.zeros((1, 20), complex)
    assert lib.chb_test_fft_lines(20, 1, 1, x.ctypes.data) != 0
    assert b"2^a" in lib.chb_last_error()
