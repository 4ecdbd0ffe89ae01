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


def test_select_tracks_host_side_frame_and_conventions(monkeypatch, tmp_path):
    """The host half of the drop-in (ysmr_b200/select.py) without a GPU: with the C call replaced by the fixture's answer the
    returned frame has the reference's columns, 'index' values and rows, the csv is written, and the None conventions hold."""
    import pandas as pd
    from ysmr_b200 import _lib, select
    d = np.load(os.path.join(GOLDEN, 'select_cfg1.npz'))
    rows = d['rows']
    cols = ['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']
    df = pd.DataFrame({c: rows[:, i] for i, c in enumerate(cols)})
    df['TRACK_ID'] = df['TRACK_ID'].astype(np.uint32); df['POSITION_T'] = df['POSITION_T'].astype(np.uint32)
    # what the device would answer: rows of the cleaned-up frame are the input rows whose (track, t) survive
    key = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(rows[:, 0], rows[:, 1]))}
    clean_rows = [key[(int(a), int(b))] for a, b in zip(d['clean_TRACK_ID'], d['clean_POSITION_T'])]
    clean_index = np.full(len(rows), -1, np.int32); clean_index[clean_rows] = np.arange(len(clean_rows))
    good = np.zeros(len(rows), np.uint8); good[np.asarray(clean_rows)[d['sel_index']]] = 1
    state = {'status': _lib.SEL_OK}

    def fake(track_id, t, x, y, w, h, params, device=0):
        info = np.zeros(_lib.SELECT_INFO); info[_lib.SI_STATUS] = state['status']
        info[_lib.SI_ROWS_BEFORE] = len(rows); info[_lib.SI_ROWS_AFTER] = len(clean_rows)
        info[_lib.SI_TRACKS_BEFORE] = 60; info[_lib.SI_TRACKS_AFTER] = len(d['top_kick']); info[_lib.SI_GOOD_TRACKS] = 50
        return good, clean_index, np.bincount(d['top_kick'], minlength=9).astype(np.int64), info
    monkeypatch.setattr(select, 'select_rows', fake)
    st = dict(zip([str(k) for k in d['setting_keys']], [float(v) for v in d['setting_values']]))
    st['store processed .csv file'] = True
    out = select.select_tracks(path_to_file=str(tmp_path / 'v_list.csv'), df=df, results_directory=str(tmp_path), fps=float(d['fps']),
                               frame_height=int(d['frame_height']), frame_width=int(d['frame_width']), settings=st)
    assert list(out.columns) == ['index'] + cols and (out['index'].to_numpy() == d['sel_index']).all()
    assert (out['TRACK_ID'].to_numpy() == d['sel_track']).all() and (out['POSITION_T'].to_numpy() == d['sel_t']).all()
    back = pd.read_csv(tmp_path / 'v_list_selected_data.csv')
    assert list(back.columns) == ['index'] + cols and len(back) == len(out)
    for status in (_lib.SEL_TOO_SHORT_BEFORE, _lib.SEL_TOO_SHORT_AFTER, _lib.SEL_NO_TRACKS):      # track_eval.py:599-606, 676-684, 816-819
        state['status'] = status
        assert select.select_tracks(path_to_file=str(tmp_path / 'v_list.csv'), df=df, results_directory=str(tmp_path), fps=30.0,
                                    frame_height=922, frame_width=1228, settings=dict(st)) is None
    state['status'] = _lib.SEL_OK
    bad = dict(st); bad['pixel per micrometre'] = 0
    assert select.select_tracks(path_to_file='x_list.csv', df=df, results_directory=str(tmp_path), fps=30.0, frame_height=922,
                                frame_width=1228, settings=bad) is None
    assert select.select_tracks(path_to_file='x_list.csv', df=df, results_directory=str(tmp_path), fps=0, frame_height=922,
                                frame_width=1228, settings=dict(st, **{'frames per second': 0})) is None
