"""CPU: the YSMR_HD logic of csrc/select.cuh (host emulation) against numpy / pandas and against the fixtures the
reference's own select_tracks / find_good_tracks produced (oracle/make_golden_select.py)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'select_*.npz')))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_pairwise_sum_is_numpy_sum_bit_for_bit(emul):
    emul.emul_np_sum.restype = C.c_double
    rng = np.random.default_rng(0)
    for n in list(range(1, 300)) + [1000, 1023, 1024, 1025, 4097, 8191, 8192, 8193, 9000, 20000, 65537, 100001]:
        a = rng.normal(100, 30, n)
        assert emul.emul_np_sum(_p(a), C.c_int64(n)) == a.sum(), n


def test_mean_is_pandas_mean_on_frame_slices(emul):
    import pandas as pd
    emul.emul_np_mean.restype = C.c_double
    rng = np.random.default_rng(1)
    df = pd.DataFrame({'a': rng.normal(10, 3, 30000), 'b': rng.normal(500, 200, 30000)})
    for s, e in [(0, 29999), (5, 4000), (123, 20011), (7, 7), (100, 101), (17, 17 + 599)]:
        v = np.ascontiguousarray(df.iloc[s:e + 1]['b'].to_numpy())
        assert emul.emul_np_mean(_p(v), C.c_int64(len(v))) == df.iloc[s:e + 1]['b'].mean()


def _cfg(d):
    st = dict(zip([str(k) for k in d['setting_keys']], d['setting_values']))
    fps = float(d['fps'])
    return st, np.array([
        int(round(fps, 0) * st['minimal length in seconds']), st['maximal consecutive holes'], st['maximal recursion depth'],
        st['maximal empty frames in %'], d['bounds'][0], d['bounds'][1], st['average width/height ratio min.'],
        st['average width/height ratio max.'], st['percent of screen edges to exclude'], float(d['frame_height']),
        float(d['frame_width']), int(round(fps, 0) * st['limit track length to x seconds']), st['limit track length exactly']], np.float64)


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_find_good_tracks_equals_reference(emul, path):
    d = np.load(path)
    _, cfg = _cfg(d)
    starts = np.ascontiguousarray(d['top_start'], np.int32)
    n_tr, n = len(starts), len(d['clean_area'])
    t = np.ascontiguousarray(d['clean_POSITION_T'], np.uint32)
    cols = [np.ascontiguousarray(d['clean_' + k], np.float64) for k in ('POSITION_X', 'POSITION_Y', 'area', 'ratio_wh')]
    outl = np.ascontiguousarray(d['clean_distance'], np.int8)
    gs, ge, kick = np.empty(n_tr, np.int32), np.empty(n_tr, np.int32), np.empty(n_tr, np.int32)
    emul.emul_select_tracks(_p(starts), n_tr, n, _p(t), _p(cols[0]), _p(cols[1]), _p(cols[2]), _p(cols[3]), _p(outl), _p(cfg),
                            _p(gs), _p(ge), _p(kick))
    assert (kick == d['top_kick']).all()
    # the rows the reference finally selected: per track one contiguous range of cleaned-up indices
    want = np.zeros(n, bool)
    want[d['sel_index']] = True
    got = np.zeros(n, bool)
    for a, b in zip(gs, ge):
        if a >= 0:
            got[a:b + 1] = True
    assert (got == want).all()
    # without a limit the chosen fragment is the longest accepted one, the first among equals
    if cfg[11] == 0:
        assert (np.stack([gs, ge], 1) == d['top_best']).all()


def test_select_params_follow_the_reference_conversions():
    from ysmr_b200.select import SELECT_DEFAULTS, select_params
    st = dict(SELECT_DEFAULTS)
    p = select_params(st, 29.97, 922, 1228)
    assert p.min_len_frames == int(round(29.97, 0) * 20.0) == 600 and p.limit_frames == 600
    assert p.max_empty == 1.05 and p.q_area == 0.1 and p.edge == 0.05 and p.max_recursion == 960
    st['exclude measurement when above x times average area'] = 0
    assert select_params(st, 30.0, 10, 10).area_factor == 0.0
