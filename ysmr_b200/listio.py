"""The per-track CSV of the hot loop: ``<video>_list.csv``.

Same text, same two-step life cycle as the reference: rows are appended as text while tracking
(helper_file.save_list, 1403-1478), then the file is read back with pandas, sorted by (TRACK_ID, POSITION_T) and
rewritten (helper_file.sort_list 1538-1574 -> get_data 846-919 -> save_df_to_csv 1366-1400).  The pandas round trip is
kept on purpose: it is what turns the '0' of an unmatched track into '0.0' and what defines the float formatting of the
final file.
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

HEADER = 'TRACK_ID,POSITION_T,POSITION_X,POSITION_Y,WIDTH,HEIGHT,DEGREES_ANGLE\n'
DTYPES = {'TRACK_ID': np.uint32, 'POSITION_T': np.uint32, 'POSITION_X': np.float64, 'POSITION_Y': np.float64,
          'WIDTH': np.float64, 'HEIGHT': np.float64, 'DEGREES_ANGLE': np.float64}


def start_list(video_path, result_folder=None, rename_old_list=True):
    """First call of save_list: decides the csv path, moves or removes an older one, writes the header.
    Returns (old_list_path_or_False, csv_path)."""
    folder, name_ext = os.path.split(video_path)
    folder = folder if result_folder is None else result_folder
    stem = os.path.splitext(name_ext)[0]
    csv_path = os.path.join(folder, '{}_list.csv'.format(stem))
    now = datetime.now().strftime('%y%m%d%H%M%S')
    old_list, denied = False, False
    if os.path.isfile(csv_path):
        if rename_old_list:
            base, ext = os.path.splitext(csv_path)
            old_list = '{}_{}{}'.format(base, now, ext)
            try:
                os.rename(csv_path, old_list)
            except PermissionError:
                denied = True
        else:
            try:
                os.remove(csv_path)
            except PermissionError:
                denied = True
    if denied:
        old_list = csv_path
        csv_path = '{}/{}_{}_list.csv'.format(folder, now, stem)
    with open(csv_path, 'w+', newline='') as fh:
        fh.write(HEADER)
    return old_list, csv_path


def reset_list(csv_path):
    """Back to the state right after start_list (header only)."""
    with open(csv_path, 'w+', newline='') as fh:
        fh.write(HEADER)


def append_rows(csv_path, rows):
    """rows: structured array (api.ROW_DTYPE) in emission order.  One text line per row, formatted like
    '{0},{1},{2},{3},{4},{5},{6}'.format(int(id), int(frame), x, y, w, h, deg): x, y are float64 (repr), w/h/deg are the
    float32 results of cv2.minAreaRect widened to Python floats, or the ints 0,0,0 of an unmatched track."""
    if len(rows) == 0:
        return
    with open(csv_path, 'a', newline='') as fh:
        fh.write(format_rows(rows))


def format_rows(rows):
    """The text append_rows writes for these rows."""
    ids = rows['track_id'].tolist(); frames = rows['frame'].tolist()
    xs = rows['x'].tolist(); ys = rows['y'].tolist()
    ws = rows['w'].astype(np.float64).tolist(); hs = rows['h'].astype(np.float64).tolist()
    ds = rows['deg'].astype(np.float64).tolist()
    out = []
    for i, t, x, y, w, h, d in zip(ids, frames, xs, ys, ws, hs, ds):
        if w == 0.0 and h == 0.0 and d == 0.0:
            out.append('{},{},{!r},{!r},0,0,0\n'.format(i, t, x, y))
        else:
            out.append('{},{},{!r},{!r},{!r},{!r},{!r}\n'.format(i, t, x, y, w, h, d))
    return ''.join(out)


def sort_list(csv_path, save_file=True):
    """Read back, sort by (TRACK_ID, POSITION_T), rewrite -- the same pandas calls as the reference."""
    import pandas as pd
    with open(csv_path, 'r', newline='\n') as fh:
        df = pd.read_csv(fh, sep=',', header=0, usecols=list(DTYPES.keys()), dtype=DTYPES)
    df.sort_values(by=['TRACK_ID', 'POSITION_T'], inplace=True, na_position='first')
    df.reset_index(drop=True, inplace=True)
    if save_file:
        with open(csv_path, 'w+', newline='\n') as fh:
            df.to_csv(fh, index=False, encoding='utf-8')
    return df


def write_sorted(csv_path, rows, save_file=True, presorted=False):
    """Row sink with a single write: the rows of the whole video (api.ROW_DTYPE; emission order, or -- with
    presorted=True -- already grouped by (track_id, frame) on the device, Context.rows_sorted, so that no host sort runs) are
    formatted as the hot loop would have appended them, parsed back IN MEMORY by the same pandas call as
    helper_file.get_data, sorted by (TRACK_ID, POSITION_T) and written once.  The parse cannot be skipped: pandas' default
    float parser is not round-trip exact (10.494321823120117 comes back as 10.494321823120115), and that last-digit
    perturbation is part of the reference's file.  tests/test_listio.py compares the bytes with append_rows + sort_list."""
    import io
    import pandas as pd
    df = pd.read_csv(io.StringIO(HEADER + format_rows(rows)), sep=',', header=0, usecols=list(DTYPES.keys()), dtype=DTYPES)
    if not presorted:
        df.sort_values(by=['TRACK_ID', 'POSITION_T'], inplace=True, na_position='first')
        df.reset_index(drop=True, inplace=True)
    if save_file:
        with open(csv_path, 'w+', newline='\n') as fh:
            df.to_csv(fh, index=False, encoding='utf-8')
    return df
