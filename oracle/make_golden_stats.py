"""ORACLE tooling: golden fixtures for the per-track statistics (partial) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden_stats``

For the cfg1 rows and one synthetic table (tests/golden/select_*.npz) the reference's select_tracks picks the rows and its
evaluate_tracks (track_eval.py:846-1318, plots switched off) computes df_stats; the selected rows and the eight columns the
GPU path reproduces are stored as tests/golden/stats_*.npz.
"""
from __future__ import annotations

import logging
import os
import tempfile

import numpy as np

from .make_golden import GOLDEN, import_reference, reference_settings

COLUMNS = ['Distance (µm)', 'Speed (µm/s)', 'Time (s)', 'Displacement (µm)', 'Perc. Motile', 'Arc-Chord Ratio', 'Bacteria Length',
           'Displacement divided by length']


def main():
    import pandas as pd
    helper_file, track_eval, _ = import_reference()
    logging.disable(logging.CRITICAL)
    cols = ['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']
    for name, over in (('cfg1', {}), ('synth_no_limit_no_quant', {'limit track length to x seconds': 0.0, 'percent quantiles excluded area': 0.0,
                                                                 'try to omit motility outliers': False})):
        d = np.load(os.path.join(GOLDEN, 'select_{}.npz'.format(name)))
        rows = d['rows']
        df = pd.DataFrame({c: rows[:, i] for i, c in enumerate(cols)})
        df['TRACK_ID'] = df['TRACK_ID'].astype(np.uint32); df['POSITION_T'] = df['POSITION_T'].astype(np.uint32)
        with tempfile.TemporaryDirectory() as tmp:
            st = reference_settings(helper_file, tmp)
            st.update(over)
            for k in list(st):
                if (k.startswith('save ') or 'plot' in k) and isinstance(st[k], bool):
                    st[k] = False
            st['store processed .csv file'] = False; st['store generated statistical .csv file'] = False
            st['store final analysed .csv file'] = False
            fps = float(d['fps'])
            sel = track_eval.select_tracks(path_to_file=os.path.join(tmp, 'a_list.csv'), df=df.copy(), results_directory=tmp, fps=fps,
                                           frame_height=int(d['frame_height']), frame_width=int(d['frame_width']), settings=dict(st))
            _, stats = track_eval.evaluate_tracks(path_to_file=os.path.join(tmp, 'a_list.csv'), results_directory=tmp, df=sel.copy(),
                                                  settings=dict(st), fps=fps)
        out = {'selected': sel[cols].to_numpy(np.float64), 'fps': np.float64(fps), 'px': np.float64(st['pixel per micrometre']),
               'track_id': stats['TRACK_ID'].to_numpy().astype(np.int64),
               'stats': np.stack([stats[c].to_numpy().astype(np.float64) for c in COLUMNS], 1)}
        np.savez_compressed(os.path.join(GOLDEN, 'stats_{}.npz'.format(name)), **out)
        print('stats_{}: {} selected rows, {} tracks'.format(name, len(sel), len(stats)))


if __name__ == '__main__':
    main()
