"""CPU: the per-track statistics logic of csrc/stats.cuh (host emulation) against the fixtures the reference's own
evaluate_tracks produced (oracle/make_golden_stats.py): all eight columns bit for bit."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'stats_*.npz')))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_statistics_equal_reference(emul, path):
    from ysmr_b200.evaluate import median_kernel
    d = np.load(path)
    sel = d['selected']
    tid = sel[:, 0].astype(np.uint32)
    starts = np.ascontiguousarray(np.flatnonzero(np.r_[True, tid[1:] != tid[:-1]]), np.int32)
    t = np.ascontiguousarray(sel[:, 1], np.uint32)
    x, y, w, h = [np.ascontiguousarray(sel[:, k], np.float64) for k in (2, 3, 4, 5)]
    out = np.empty((len(starts), 8), np.float64)
    emul.emul_track_statistics(_p(starts), len(starts), len(tid), _p(t), _p(x), _p(y), _p(w), _p(h), C.c_double(float(d['px'])),
                               C.c_double(float(d['fps'])), median_kernel(float(d['fps'])), _p(out))
    assert (tid[starts] == d['track_id']).all()
    assert (out == d['stats']).all(), np.abs(out - d['stats']).max(0)


def test_half_rounding_is_numpy_float16(emul):
    """f16_round (stats.cuh) through the bacteria-length mean: a one-row track returns float32(float16(max(w, h) / px))."""
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.uniform(0.01, 40.0, 400), [2049.0, 2050.0, 2051.0, 5.0e-8, 6.1e-5, 65519.0]])
    for v in vals:
        t = np.zeros(1, np.uint32); z = np.zeros(1); w = np.array([v]); h = np.array([v * 0.5])
        out = np.empty((1, 8)); starts = np.zeros(1, np.int32)
        emul.emul_track_statistics(_p(starts), 1, 1, _p(t), _p(z), _p(z), _p(w), _p(h), C.c_double(1.0), C.c_double(30.0), 31, _p(out))
        assert out[0, 6] == float(np.float32(np.float16(v))), v
