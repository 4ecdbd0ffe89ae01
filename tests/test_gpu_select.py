"""GPU: the select_tracks drop-in (ysmr_b200/select.py -> ysmr_select_tracks, csrc/select.cu) against the fixtures written by
the reference's own select_tracks (oracle/make_golden_select.py): same rows, same 'index' column, same kick reasons."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'select_*.npz')))
COLS = ['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']


def _frame(rows):
    import pandas as pd
    df = pd.DataFrame({c: rows[:, i] for i, c in enumerate(COLS)})
    df['TRACK_ID'] = df['TRACK_ID'].astype(np.uint32)
    df['POSITION_T'] = df['POSITION_T'].astype(np.uint32)
    return df


def _settings(d):
    st = dict(zip([str(k) for k in d['setting_keys']], [float(v) for v in d['setting_values']]))
    for k in ('limit track length exactly', 'try to omit motility outliers'):
        st[k] = bool(st[k])
    for k in ('maximal consecutive holes', 'maximal recursion depth'):
        st[k] = int(st[k])
    st['store processed .csv file'] = False
    return st


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_select_tracks_equals_reference(path, tmp_path):
    from ysmr_b200.select import select_tracks
    d = np.load(path)
    out = select_tracks(path_to_file=str(tmp_path / 'x_list.csv'), df=_frame(d['rows']), results_directory=str(tmp_path),
                        fps=float(d['fps']), frame_height=int(d['frame_height']), frame_width=int(d['frame_width']),
                        settings=_settings(d))
    assert (out is None) == bool(d['returned_none'])
    assert list(out.columns) == ['index'] + COLS
    assert (out['index'].to_numpy() == d['sel_index']).all()
    assert (out['TRACK_ID'].to_numpy() == d['sel_track']).all() and (out['POSITION_T'].to_numpy() == d['sel_t']).all()
    assert out.attrs['kick_reasons'] == np.bincount(d['top_kick'], minlength=9).tolist()
    # the other columns are the input rows, untouched
    src = d['rows']
    key = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(src[:, 0], src[:, 1]))}
    take = [key[(int(a), int(b))] for a, b in zip(out['TRACK_ID'], out['POSITION_T'])]
    assert (out[COLS[2:]].to_numpy() == src[take][:, 2:]).all()


def test_select_writes_csv_and_none_conventions(tmp_path):
    import pandas as pd
    from ysmr_b200.select import select_tracks
    d = np.load(os.path.join(GOLDEN, 'select_cfg1.npz'))
    st = _settings(d)
    st['store processed .csv file'] = True
    df = _frame(d['rows'])
    csv = tmp_path / 'vid_list.csv'
    df.to_csv(csv, index=False)
    out = select_tracks(path_to_file=str(csv), results_directory=str(tmp_path), fps=float(d['fps']),
                        frame_height=int(d['frame_height']), frame_width=int(d['frame_width']), settings=st)
    back = pd.read_csv(tmp_path / 'vid_list_selected_data.csv')
    assert len(back) == len(out) == len(d['sel_index']) and list(back.columns) == ['index'] + COLS
    # too few rows for the minimal length -> None, like the reference (track_eval.py:599-606)
    st2 = dict(st); st2['minimal length in seconds'] = 1.0e4
    assert select_tracks(path_to_file=str(csv), df=df, results_directory=str(tmp_path), fps=30.0, frame_height=922,
                         frame_width=1228, settings=st2) is None
    # area limits the wrong way round -> None (track_eval.py:588-597)
    st3 = dict(st); st3['extreme area outliers lower end in px*px'] = 60
    assert select_tracks(path_to_file=str(csv), df=df, results_directory=str(tmp_path), fps=30.0, frame_height=922,
                         frame_width=1228, settings=st3) is None


def test_select_scales_to_a_long_video():
    """2.7 M rows (the row count of the 54,000-frame cfg5 video): per-track invariants that do not need the reference --
    every selected fragment is one contiguous run of one track, no longer than the limit, and re-selecting the selected
    rows keeps all of them."""
    from ysmr_b200 import _lib
    from ysmr_b200.select import SELECT_DEFAULTS, select_params, select_rows
    rng = np.random.default_rng(5)
    n_tracks, length = 300, 9000
    tid = np.repeat(np.arange(n_tracks, dtype=np.uint32), length)
    t = np.tile(np.arange(length, dtype=np.uint32), n_tracks)
    x = np.repeat(rng.uniform(200, 1000, n_tracks), length) + rng.normal(0, 0.3, n_tracks * length).cumsum() * 0.01
    y = np.repeat(rng.uniform(200, 700, n_tracks), length) + rng.normal(0, 0.2, n_tracks * length)
    w = rng.uniform(7.5, 9.5, n_tracks * length).astype(np.float32).astype(np.float64)
    h = rng.uniform(2.2, 3.2, n_tracks * length).astype(np.float32).astype(np.float64)
    p = select_params(dict(SELECT_DEFAULTS), 30.0, 922, 1228)
    good, clean, kicks, info = select_rows(tid, t, x, y, w, h, p)
    assert int(info[_lib.SI_STATUS]) == 0 and kicks.sum() == int(info[_lib.SI_TRACKS_AFTER])
    sel = np.flatnonzero(good)
    assert len(sel) > 0
    runs = np.split(sel, np.flatnonzero(np.diff(sel) != 1) + 1)
    assert len(runs) == int(info[_lib.SI_GOOD_TRACKS])
    for r in runs:
        assert len(np.unique(tid[r])) == 1 and t[r[-1]] - t[r[0]] + 1 <= p.limit_frames


def test_pipeline_selection_equals_the_reference_chain():
    """frames -> GPU detect + link -> device-sorted rows -> GPU select_tracks, against
    frames -> reference track_bacteria -> reference select_tracks (fixtures e2e_cfg1_300 / select_cfg1): the same tracks and
    frames are selected, the selected rows carry the same values to the parity bars of the hot path."""
    import pandas as pd
    torch = pytest.importorskip('torch')
    from ysmr_b200.api import Context
    from ysmr_b200.select import select_tracks
    from ysmr_b200.synth import SceneConfig, make_scene, render_frames
    g = np.load(os.path.join(GOLDEN, 'e2e_cfg1_300.npz'))
    d = np.load(os.path.join(GOLDEN, 'select_cfg1.npz'))
    cfg = SceneConfig(**{k[6:]: g[k].item() for k in g.files if k.startswith('scene_')})
    grey = render_frames(make_scene(cfg))
    ctx = Context(cfg.height, cfg.width, 1, 0, max_batch=64, max_blobs=1024, max_tracks=1024)
    ctx.archive_rows(True)
    ctx.track_host(grey, 0, copy_rows=False)
    rows = ctx.rows_sorted()                                   # (TRACK_ID, POSITION_T) order, from the device
    ctx.close()
    df = pd.DataFrame({'TRACK_ID': rows['track_id'].astype(np.uint32), 'POSITION_T': rows['frame'].astype(np.uint32),
                       'POSITION_X': rows['x'], 'POSITION_Y': rows['y'], 'WIDTH': rows['w'].astype(np.float64),
                       'HEIGHT': rows['h'].astype(np.float64), 'DEGREES_ANGLE': rows['deg'].astype(np.float64)})
    out = select_tracks(path_to_file='cfg1_list.csv', df=df, results_directory='.', fps=float(d['fps']),
                        frame_height=int(d['frame_height']), frame_width=int(d['frame_width']), settings=_settings(d))
    assert out is not None
    assert (out['TRACK_ID'].to_numpy() == d['sel_track']).all() and (out['POSITION_T'].to_numpy() == d['sel_t']).all()
    assert (out['index'].to_numpy() == d['sel_index']).all()
    ref = d['rows']
    key = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(ref[:, 0], ref[:, 1]))}
    take = [key[(int(a), int(b))] for a, b in zip(out['TRACK_ID'], out['POSITION_T'])]
    assert np.abs(out[['POSITION_X', 'POSITION_Y']].to_numpy() - ref[take][:, 2:4]).max() < 1e-3
    assert np.abs(out[['WIDTH', 'HEIGHT']].to_numpy() - ref[take][:, 4:6]).max() < 1e-3
