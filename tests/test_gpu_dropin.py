"""-m gpu: the drop-in ``track_bacteria`` (ysmr_b200/track_eval.py) on a lossless FFV1 AVI against the golden rows and
CSV text of the reference's track_bacteria (track_eval.py:38-405) on the same bytes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from tests.util import GOLDEN  # noqa: E402
from ysmr_b200.synth import SceneConfig, make_scene, render_frames  # noqa: E402


def _write_ffv1(path, grey, fps):
    import cv2
    h, w = grey.shape[1:]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'FFV1'), fps, (w, h), True)
    assert vw.isOpened()
    for f in grey:
        vw.write(np.repeat(f[..., None], 3, axis=-1))
    vw.release()


@pytest.mark.parametrize('name,sink', [('small_wod', 'append'), ('small_dol', 'append'), ('small_wod', 'once')])
def test_track_bacteria_dropin(tmp_path, name, sink):
    from ysmr_b200.track_eval import track_bacteria
    g = np.load(os.path.join(GOLDEN, f'e2e_{name}.npz'))
    kw = {k[6:]: g[k].item() for k in g.files if k.startswith('scene_')}
    cfg = SceneConfig(**kw)
    grey = render_frames(make_scene(cfg))
    video = str(tmp_path / f'{name}.avi')
    _write_ffv1(video, grey, cfg.fps)
    settings = {'white bacteria on dark background': bool(g['white_on_dark']), 'threshold offset for detection': int(g['offset']),
                'adaptive double threshold': float(g['adt']), 'minimal frame count': 10, 'display video analysis': False}
    res = track_bacteria(video, settings, str(tmp_path), chunk_frames=64, max_blobs=512, max_tracks=512, row_sink=sink)
    assert res is not None
    df, fps, fh, fw, csv = res
    assert (fps, fh, fw) == (float(g['fps']), int(g['frame_height']), int(g['frame_width']))
    assert os.path.basename(csv) == f'{name}_list.csv'
    # the reference flips the sign of the caller's offset in place for dark-on-light (track_eval.py:132); so do we
    assert settings['threshold offset for detection'] == (int(g['offset']) if bool(g['white_on_dark']) else -int(g['offset']))
    ref = g['rows']
    mine = df[['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']].to_numpy(np.float64)
    assert mine.shape == ref.shape and (mine[:, :2] == ref[:, :2]).all()
    assert list(df.dtypes.astype(str)) == ['uint32', 'uint32', 'float64', 'float64', 'float64', 'float64', 'float64']
    assert np.abs(mine[:, 4:] - ref[:, 4:]).max() < 1e-3
    head = open(csv).readlines()[0]
    assert head == 'TRACK_ID,POSITION_T,POSITION_X,POSITION_Y,WIDTH,HEIGHT,DEGREES_ANGLE\n'
    # same text layout as the reference's file (ints, then repr floats); values agree to the tolerances above
    mine_lines = open(csv).readlines()[1:4]
    ref_lines = str(g['csv_head']).splitlines()[1:4]
    for a, b in zip(mine_lines, ref_lines):
        fa, fb = a.strip().split(','), b.strip().split(',')
        assert fa[:2] == fb[:2] and len(fa) == len(fb) == 7
        assert np.allclose([float(v) for v in fa[2:]], [float(v) for v in fb[2:]], rtol=1e-9, atol=1e-9)
        assert fa[4:] == fb[4:]                       # w, h, deg: identical float32 values -> identical text


def test_track_bacteria_error_conventions(tmp_path):
    from ysmr_b200.track_eval import track_bacteria
    assert track_bacteria(str(tmp_path / 'missing.avi'), {'minimal frame count': 1}, str(tmp_path)) is None
    grey = render_frames(make_scene(SceneConfig(width=64, height=48, n_frames=5, n_cells=1, seed=1, margin=10.0)))
    video = str(tmp_path / 'short.avi')
    _write_ffv1(video, grey, 30.0)
    assert track_bacteria(video, {'minimal frame count': 600}, str(tmp_path)) is None            # too short: skipped
    assert track_bacteria(video, {'minimal frame count': 1, 'include luminosity in tracking calculation': True},
                          str(tmp_path)) is None
