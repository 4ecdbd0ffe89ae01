"""ORACLE tooling: golden fixtures for the track selection from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden_select``

Runs ``ysmr.track_eval.select_tracks`` (track_eval.py:541-843) on (a) the rows the reference's own track_bacteria produced
for the cfg1 video (tests/golden/e2e_cfg1_300.npz) and (b) seeded synthetic track tables with frame gaps, position jumps,
unmatched rows, out-of-range areas and tracks near the frame edge, under several settings.  ('limit track length exactly' has
no fixture: with pandas 3 the reference raises "Encountered all NA values" at track_eval.py:788 as soon as one fragment has
no row at exactly the limit; the GPU path keeps the older pandas behaviour the code was written for -- such a track is
skipped, :791-792.)  ``find_good_tracks`` is
wrapped (not modified) to record its top-level results and the cleaned-up frame it works on.  Writes
tests/golden/select_*.npz.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

from .make_golden import GOLDEN, import_reference, reference_settings

SELECT_KEYS = [
    'minimal length in seconds', 'limit track length to x seconds', 'limit track length exactly',
    'extreme area outliers lower end in px*px', 'extreme area outliers upper end in px*px',
    'exclude measurement when above x times average area', 'maximal consecutive holes', 'maximal empty frames in %',
    'percent quantiles excluded area', 'try to omit motility outliers',
    'stop excluding motility outliers if total count above percent', 'average width/height ratio min.',
    'average width/height ratio max.', 'percent of screen edges to exclude', 'maximal recursion depth',
]


def synthetic_table(seed, n_tracks=36, frame_h=922, frame_w=1228):
    """A sorted track table with the defects select_tracks has to deal with."""
    rng = np.random.default_rng(seed)
    rows = []
    for tid in range(n_tracks):
        length = int(rng.integers(20, 500))
        t0 = int(rng.integers(0, 300))
        t = t0 + np.arange(length)
        # frame gaps: drop single frames, short runs, and sometimes a long run (> 'maximal consecutive holes')
        keep = np.ones(length, bool)
        for _ in range(int(rng.integers(0, 4))):
            a = int(rng.integers(1, max(2, length - 1)))
            keep[a:a + int(rng.choice([1, 1, 2, 3, 8, 15]))] = False
        keep[0] = True
        t = t[keep]
        n = len(t)
        edge = rng.random() < 0.15
        x0 = rng.uniform(0, 40) if edge else rng.uniform(150, frame_w - 150)
        y0 = rng.uniform(150, frame_h - 150)
        v = rng.normal(0, 0.4, (n, 2)).cumsum(0) * 0.3
        x = x0 + v[:, 0] + rng.normal(0, 0.15, n)
        y = y0 + v[:, 1] + rng.normal(0, 0.15, n)
        for _ in range(int(rng.integers(0, 3))):                      # position jumps (motility outliers)
            if rng.random() < 0.5 and n > 10:
                k = int(rng.integers(5, n - 1))
                x[k:] += rng.uniform(20, 60)
        big = rng.random() < 0.1                                      # median area outside [2, 50]
        w = np.float32(rng.uniform(60, 90) if big else rng.uniform(7.5, 9.5)) + rng.normal(0, 0.3, n).astype(np.float32)
        h = np.float32(rng.uniform(2.2, 3.2)) + rng.normal(0, 0.15, n).astype(np.float32)
        if rng.random() < 0.15:                                       # roundish: ratio outside the rod preset
            h = w * np.float32(0.9)
        lost = rng.random(n) < 0.008                                  # unmatched rows: w = h = deg = 0
        w = np.where(lost, np.float32(0), w); h = np.where(lost, np.float32(0), h)
        spike = rng.random(n) < 0.006                                 # merged blobs: area > 1.5 x median
        w = np.where(spike & ~lost, w * np.float32(2.2), w)
        deg = np.where(lost, 0.0, rng.uniform(-90, 0, n))
        for k in range(n):
            rows.append((tid, int(t[k]), float(x[k]), float(y[k]), float(np.float32(w[k])), float(np.float32(h[k])), float(deg[k])))
    return np.array(rows, np.float64)


def run_case(track_eval, settings, rows, fps, frame_h, frame_w, name):
    import pandas as pd
    cols = ['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']
    df = pd.DataFrame({c: rows[:, i] for i, c in enumerate(cols)})
    df['TRACK_ID'] = df['TRACK_ID'].astype(np.uint32)
    df['POSITION_T'] = df['POSITION_T'].astype(np.uint32)
    captured = {'top': [], 'df': None, 'bounds': None}
    original = track_eval.find_good_tracks

    def wrapper(*args, **kwargs):
        res = original(*args, **kwargs)
        if kwargs.get('recursion', 0) == 0:
            captured['top'].append((kwargs['start'], kwargs['stop'], res[1], list(res[0])))
            if captured['df'] is None:
                d = kwargs['df_passed']
                captured['df'] = {k: d[k].to_numpy().copy() for k in
                                  ('TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'area', 'ratio_wh', 'distance')}
                captured['bounds'] = (float(kwargs['lower_boundary']), float(kwargs['upper_boundary']))
        return res

    track_eval.find_good_tracks = wrapper
    try:
        with tempfile.TemporaryDirectory() as tmp:
            st = dict(settings)
            st['store processed .csv file'] = False
            out = track_eval.select_tracks(path_to_file=os.path.join(tmp, name + '_list.csv'), df=df.copy(), results_directory=tmp,
                                           fps=fps, frame_height=frame_h, frame_width=frame_w, settings=st)
    finally:
        track_eval.find_good_tracks = original
    data = {
        'rows': rows, 'fps': np.float64(fps), 'frame_height': np.int64(frame_h), 'frame_width': np.int64(frame_w),
        'setting_keys': np.array(SELECT_KEYS), 'setting_values': np.array([float(settings[k]) for k in SELECT_KEYS], np.float64),
        'returned_none': np.bool_(out is None),
    }
    if out is not None:
        data['sel_index'] = out['index'].to_numpy().astype(np.int64)
        data['sel_track'] = out['TRACK_ID'].to_numpy().astype(np.int64)
        data['sel_t'] = out['POSITION_T'].to_numpy().astype(np.int64)
    if captured['df'] is not None:
        for k, v in captured['df'].items():
            data['clean_' + k] = v
        data['bounds'] = np.array(captured['bounds'], np.float64)
        data['top_start'] = np.array([a for a, _, _, _ in captured['top']], np.int64)
        data['top_stop'] = np.array([b for _, b, _, _ in captured['top']], np.int64)
        data['top_kick'] = np.array([k for _, _, k, _ in captured['top']], np.int64)
        # the longest fragment, the first among equals (track_eval.py:771-779), before the length limit
        best = []
        for _, _, _, frags in captured['top']:
            if not frags:
                best.append((-1, -1))
            else:
                lens = [b - a + 1 for a, b in frags]
                best.append(frags[int(np.argmax(lens))])
        data['top_best'] = np.array(best, np.int64)
    np.savez_compressed(os.path.join(GOLDEN, 'select_{}.npz'.format(name)), **data)
    n_sel = 0 if out is None else len(out)
    print('select_{}: rows {} -> selected {} rows, {} tracks, kicks {}'.format(
        name, len(rows), n_sel, 0 if out is None else out['TRACK_ID'].nunique(),
        np.bincount(data.get('top_kick', np.zeros(0, np.int64)), minlength=9).tolist()))


def main():
    helper_file, track_eval, _ = import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        base = reference_settings(helper_file, tmp)
    base['verbose'] = False
    e2e = np.load(os.path.join(GOLDEN, 'e2e_cfg1_300.npz'))
    run_case(track_eval, base, e2e['rows'], float(e2e['fps']), int(e2e['frame_height']), int(e2e['frame_width']), 'cfg1')
    variants = {
        'synth_default': {},
        'synth_no_limit_no_quant': {'limit track length to x seconds': 0.0, 'percent quantiles excluded area': 0.0,
                                    'try to omit motility outliers': False},
        'synth_shallow': {'maximal recursion depth': 1, 'maximal consecutive holes': 2, 'minimal length in seconds': 1.0},
        'synth_many_outliers': {'stop excluding motility outliers if total count above percent': 0.0005},
    }
    for i, (name, over) in enumerate(variants.items()):
        st = dict(base)
        st.update(over)
        run_case(track_eval, st, synthetic_table(100 + i), 30.0, 922, 1228, name)


if __name__ == '__main__':
    main()
