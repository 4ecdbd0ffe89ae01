"""Row sinks of the drop-in (ysmr_b200/listio.py): the reference's append-then-sort life cycle (helper_file.py:1403-1478,
1538-1574) against the single sorted write -- same bytes, same DataFrame -- and the reason the text parse
cannot be skipped (pandas' default float parser perturbs the last digit)."""
import os

import numpy as np
import pandas as pd

from ysmr_b200 import listio
from ysmr_b200.api import ROW_DTYPE


def _rows(seed, n_frames=120, n_tracks=23):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n_frames):
        ids = np.flatnonzero(rng.random(n_tracks) < 0.8)
        rng.shuffle(ids)                                            # emission order inside a frame is not sorted
        for i in ids:
            r = np.zeros((), ROW_DTYPE)
            r['frame'], r['track_id'] = t, i
            r['x'], r['y'] = rng.uniform(0, 1228), rng.uniform(0, 922)
            if rng.random() < 0.85:                                 # matched: float32 sizes from minAreaRect
                r['w'], r['h'], r['deg'] = np.float32(rng.uniform(1, 12)), np.float32(rng.uniform(1, 12)), np.float32(-rng.uniform(0, 90))
            out.append(r)
    return np.array(out, ROW_DTYPE)


def test_single_sorted_write_is_byte_identical(tmp_path):
    for seed in range(3):
        rows = _rows(seed)
        a = str(tmp_path / f'a{seed}_list.csv'); b = str(tmp_path / f'b{seed}_list.csv')
        with open(a, 'w', newline='') as fh:
            fh.write(listio.HEADER)
        for k in range(0, len(rows), 777):                          # appended in pieces like the hot loop does
            listio.append_rows(a, rows[k:k + 777])
        df_a = listio.sort_list(a)
        df_b = listio.write_sorted(b, rows)
        assert open(a, 'rb').read() == open(b, 'rb').read()
        pd.testing.assert_frame_equal(df_a, df_b)
        assert list(df_b.dtypes) == [np.uint32, np.uint32] + [np.float64] * 5


def test_unmatched_rows_are_zero_floats(tmp_path):
    rows = np.zeros(2, ROW_DTYPE)
    rows['track_id'] = [1, 0]; rows['frame'] = [5, 5]; rows['x'] = [1.5, 2.25]; rows['y'] = [3.0, 4.0]
    p = str(tmp_path / 'z_list.csv')
    df = listio.write_sorted(p, rows)
    assert open(p).read().splitlines()[1] == '0,5,2.25,4.0,0.0,0.0,0.0'
    assert df['TRACK_ID'].tolist() == [0, 1]


def test_pandas_round_trip_is_not_exact_and_is_kept(tmp_path):
    # the float32 width 10.494322 widens to 10.494321823120117; the reference's read-back turns it into ...115
    rows = np.zeros(1, ROW_DTYPE)
    rows['x'], rows['y'] = 952.5487789890452, 284.76648842715866
    rows['w'], rows['h'], rows['deg'] = np.float32(10.494321823120117), np.float32(10.694378852844238), np.float32(-45.963584899902344)
    assert repr(float(rows['w'][0])) == '10.494321823120117'
    df = listio.write_sorted(str(tmp_path / 'p_list.csv'), rows)
    line = open(str(tmp_path / 'p_list.csv')).read().splitlines()[1]
    import io
    want = pd.read_csv(io.StringIO(listio.HEADER + listio.format_rows(rows)), dtype=listio.DTYPES)
    assert line.split(',')[4] == repr(float(want['WIDTH'][0]))
    assert df['WIDTH'][0] == want['WIDTH'][0]
