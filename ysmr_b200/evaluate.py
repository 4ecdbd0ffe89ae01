"""Per-track statistics of ``evaluate_tracks`` on the GPU -- PARTIAL.

``track_statistics(df, settings, fps)`` returns, for the selected rows ``select_tracks`` produced, eight of the twelve columns
of the reference's ``df_stats`` (ysmr/track_eval.py:1030-1120) bit for bit: distance, speed, time, displacement (largest
pairwise distance -- the O(L^2) part of the reference), percent motile, arc-chord ratio, bacteria length and displacement
divided by length.  Turn points, motility phenotype and median speed (track_eval.py:946-1029) are not built, so this is NOT a
drop-in for ``evaluate_tracks``: the drop-in chain (ysmr_b200/main.py) still hands the selected rows to the reference's own
``evaluate_tracks``.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

COLUMNS = ['Distance (µm)', 'Speed (µm/s)', 'Time (s)', 'Displacement (µm)', 'Perc. Motile', 'Arc-Chord Ratio', 'Bacteria Length',
           'Displacement divided by length']


def median_kernel(fps):
    """track_eval.py:933-936: round(fps), plus one if even."""
    k = int(round(fps, 0))
    return k + 1 if k & 1 == 0 else k


def track_statistics(df, settings, fps, device=0):
    import pandas as pd
    lib = _lib.load()
    tid = np.ascontiguousarray(df['TRACK_ID'].to_numpy(), np.uint32)
    cols = [tid, np.ascontiguousarray(df['POSITION_T'].to_numpy(), np.uint32)] + \
           [np.ascontiguousarray(df[k].to_numpy(), np.float64) for k in ('POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT')]
    starts = np.ascontiguousarray(np.flatnonzero(np.r_[True, tid[1:] != tid[:-1]]), np.int32)
    out = np.empty((len(starts), len(COLUMNS)), np.float64)
    rc = lib.ysmr_track_statistics(int(device), len(tid), *[a.ctypes.data_as(C.c_void_p) for a in cols],
                                   float(settings['pixel per micrometre']), float(fps), median_kernel(fps),
                                   starts.ctypes.data_as(C.c_void_p), len(starts), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError('ysmr_track_statistics failed ({}): {}'.format(rc, lib.ysmr_statistics_last_error().decode()))
    stats = pd.DataFrame(out, columns=COLUMNS, index=pd.Index(tid[starts], name='TRACK_ID'))
    stats['Bacteria Length'] = stats['Bacteria Length'].astype(np.float32)      # the reference's column is float32
    stats['TRACK_ID'] = tid[starts]
    return stats
