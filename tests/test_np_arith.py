"""The float64 roundings of NumPy that the linker restates (ysmr_b200/csrc/link.cuh: np_exp_nonpos, blas_row_dot),
pinned against the installed NumPy bit for bit.  The reference's GSFF feeds its own output back while a track is unmatched
(tracker.py:219-227), which amplifies a last-bit difference ~2.5x per frame, so "the same formula" is not enough: the
filter must round like gsff.py's numpy.dot / numpy.exp / numpy.sum calls (gsff.py:177, 193, 242, 321-337)."""
import numpy as np
import pytest

from tests.util import ptr


def _avx512():
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as f
    except Exception:                                                     # pragma: no cover
        return False
    return bool(f.get('AVX512_SKX'))


@pytest.mark.skipif(not _avx512(), reason='numpy.exp only takes the SVML path (the one restated) on AVX-512 hosts')
def test_np_exp_restatement_is_bit_identical(emul):
    rng = np.random.default_rng(5)
    x = np.concatenate([-rng.random(400000) * 47, -rng.random(100000) * 1e-2, -rng.random(100000) * 1e-7,
                        [0.0, -0.0, -46.9999, -1e-300, -46.051701859880914]])
    out = np.empty_like(x)
    emul.emul_np_exp(ptr(x), ptr(out), len(x))
    assert (out == np.exp(x)).all()
    # scalar calls (what gsff.py:193 makes) go through the same loop
    assert all(np.exp(np.float64(v)) == o for v, o in zip(x[:2000], out[:2000]))


def test_blas_row_dot_restatement_is_bit_identical(emul):
    """numpy.dot(gain (4 x 2n), list of 2n floats) as gsff.py:177 calls it, for the default horizons, odd horizons (tail
    of two positions) and the largest supported one."""
    rng = np.random.default_rng(6)
    for n in (10, 20, 30, 8, 16, 25, 7, 33, 64):
        for _ in range(100):
            g = rng.standard_normal((4, 2 * n)); y = rng.random(2 * n) * 1000
            ref = np.dot(g, y.tolist())
            for row in range(4):
                assert emul.emul_blas_row_dot(ptr(np.ascontiguousarray(g[row])), ptr(y), n) == ref[row]


def test_small_sums_and_dots_round_like_the_restatement():
    """numpy.sum(axis=1) over <= 4 products is left to right; the 2-element numpy.dot is fma(d1, d1, d0*d0) -- what
    gsff_weighted / gsff_likelihood assume."""
    import ctypes
    lm = ctypes.CDLL('libm.so.6'); lm.fma.restype = ctypes.c_double; lm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(7)
    for _ in range(3000):
        xh = rng.standard_normal((2, 3)) * 800; w = rng.random(3)
        p = xh * w
        assert ((p[:, 0] + p[:, 1]) + p[:, 2] == np.sum(xh * w, axis=1)).all()
        lik = [rng.random(), rng.random(), rng.random()]
        q = lik * w
        assert sum(q) == ((0 + q[0]) + q[1]) + q[2]
        d = rng.standard_normal(2) * rng.choice([1e-3, 1.0, 10.0])
        assert np.dot(d.T, np.dot(np.eye(2), d)) == lm.fma(d[1], d[1], d[0] * d[0])
