"""GPU: ysmr_track_statistics (csrc/stats.cu) behind ysmr_b200.evaluate.track_statistics against the reference's
evaluate_tracks (fixtures from oracle/make_golden_stats.py): eight df_stats columns bit for bit."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'stats_*.npz')))
COLS = ['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_track_statistics_equal_reference(path):
    import pandas as pd
    from ysmr_b200.evaluate import COLUMNS, track_statistics
    d = np.load(path)
    sel = d['selected']
    df = pd.DataFrame({c: sel[:, i] for i, c in enumerate(COLS)})
    df['TRACK_ID'] = df['TRACK_ID'].astype(np.uint32); df['POSITION_T'] = df['POSITION_T'].astype(np.uint32)
    stats = track_statistics(df, {'pixel per micrometre': float(d['px'])}, float(d['fps']))
    assert (stats['TRACK_ID'].to_numpy() == d['track_id']).all()
    got = stats[COLUMNS].to_numpy(np.float64)
    assert (got == d['stats']).all(), np.abs(got - d['stats']).max(0)
    assert stats['Bacteria Length'].dtype == np.float32


def test_long_tracks_displacement_is_the_pairwise_maximum():
    """9,000-row tracks (no length limit): the displacement equals a numpy evaluation of max pdist on a strided subset bound and
    the distance equals the Kahan sum of the per-row steps."""
    import pandas as pd
    from ysmr_b200.evaluate import track_statistics
    rng = np.random.default_rng(9)
    n_tr, L = 40, 9000
    tid = np.repeat(np.arange(n_tr, dtype=np.uint32), L); t = np.tile(np.arange(L, dtype=np.uint32), n_tr)
    x = rng.normal(0, 0.4, n_tr * L).cumsum() % 900 + 100; y = rng.normal(0, 0.4, n_tr * L).cumsum() % 600 + 100
    df = pd.DataFrame({'TRACK_ID': tid, 'POSITION_T': t, 'POSITION_X': x, 'POSITION_Y': y, 'WIDTH': np.full(n_tr * L, 8.0),
                       'HEIGHT': np.full(n_tr * L, 2.5), 'DEGREES_ANGLE': np.zeros(n_tr * L)})
    px = 1.41888781
    stats = track_statistics(df, {'pixel per micrometre': px}, 30.0)
    for k in (0, n_tr - 1):
        xs, ys = (x[k * L:(k + 1) * L] - x[k * L]) / px, (y[k * L:(k + 1) * L] - y[k * L]) / px
        best = 0.0
        for i in range(0, L, 1500):                       # exact rows of the distance matrix (a subset of all pairs)
            best = max(best, float(np.sqrt(np.max((xs - xs[i]) ** 2 + (ys - ys[i]) ** 2))))
        assert stats['Displacement (µm)'].iloc[k] >= best
        hull = np.sqrt((xs.max() - xs.min()) ** 2 + (ys.max() - ys.min()) ** 2)
        assert stats['Displacement (µm)'].iloc[k] <= hull
        assert stats['Time (s)'].iloc[k] == L / 30.0
