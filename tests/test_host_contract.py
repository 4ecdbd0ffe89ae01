"""Host-side contract checks that need no GPU: error conventions of the drop-in before any device work
(track_eval.py:38-155 of the reference: log + return None), and documentation / header consistency."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_track_bacteria_returns_none_before_touching_the_gpu(tmp_path):
    from ysmr_b200.track_eval import track_bacteria
    assert track_bacteria(str(tmp_path / 'missing.avi'), {'minimal frame count': 1}, str(tmp_path)) is None
    video = tmp_path / 'empty.avi'
    video.write_bytes(b'')                                   # exists, but cv2 cannot open it
    assert track_bacteria(str(video), {'minimal frame count': 1}, str(tmp_path)) is None
    assert track_bacteria(str(video), {'minimal frame count': 1}, str(tmp_path), row_sink='sometimes') is None


def test_integration_doc_lists_every_exported_symbol():
    header = open(os.path.join(ROOT, 'include', 'ysmr_b200.h')).read()
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    symbols = sorted(set(re.findall(r'\b(ysmr_[a-z_0-9]+)\s*\(', header)))
    assert len(symbols) >= 20
    missing = [s for s in symbols if s not in doc]
    assert not missing, missing


def test_header_cites_the_reference_for_every_entry_point():
    header = open(os.path.join(ROOT, 'include', 'ysmr_b200.h')).read()
    assert header.count('.py:') >= 10                        # file:line citations of the reference interface replaced


def test_horizons_follow_the_reference_operand_types():
    """gsff.py:103-106 with n_max = the float fps (tracker.py:58-59): int(fps) would give other horizons."""
    from ysmr_b200.api import horizon_sizes
    assert horizon_sizes(0, 30, 3) == [10, 20, 30] and horizon_sizes(0, 30.0, 3) == [10, 20, 30]
    assert horizon_sizes(0, 29.97, 4) == [7, 14, 22, 29] and horizon_sizes(0, 29, 4) == [7, 14, 21, 29]


def test_analyse_chain_without_the_reference_package(monkeypatch, tmp_path):
    """ysmr_b200.main.analyse when the reference package is not importable: video -> track_bacteria -> select_tracks,
    *_list.csv -> select_tracks, return conventions (DataFrame / True / None), csv deletion."""
    import pandas as pd
    from ysmr_b200 import main
    monkeypatch.setattr(main, 'install', lambda: False)
    calls = []
    csv = tmp_path / 'v_list.csv'
    csv.write_text('TRACK_ID,POSITION_T\n0,0\n')
    df_list = pd.DataFrame({'TRACK_ID': [0], 'POSITION_T': [0]})
    df_sel = pd.DataFrame({'index': [0], 'TRACK_ID': [0], 'POSITION_T': [0]})

    def fake_track(video_path, settings=None, result_folder=None):
        calls.append(('track', video_path))
        return df_list, 30.0, 922, 1228, str(csv)

    def fake_select(path_to_file=None, df=None, results_directory=None, settings=None, **meta):
        calls.append(('select', path_to_file, meta.get('fps'), df is not None))
        return df_sel
    monkeypatch.setattr(main, 'track_bacteria', fake_track)
    monkeypatch.setattr(main, 'select_tracks', fake_select)
    st = {'delete .csv file after analysis': False}
    assert main.analyse(str(tmp_path / 'v.avi'), settings=dict(st), result_folder=str(tmp_path)) is True
    assert calls == [('track', str(tmp_path / 'v.avi')), ('select', str(csv), 30.0, True)]
    out = main.analyse(str(csv), settings=dict(st), result_folder=str(tmp_path), return_df=True)
    assert out is df_sel and calls[-1] == ('select', str(csv), None, False)
    assert main.analyse(str(tmp_path / 'v_statistics.csv'), settings=dict(st)) is None
    monkeypatch.setattr(main, 'select_tracks', lambda **kw: None)
    assert main.analyse(str(csv), settings=dict(st), result_folder=str(tmp_path)) is None
    monkeypatch.setattr(main, 'select_tracks', fake_select)
    assert main.analyse(str(tmp_path / 'v.avi'), settings={'delete .csv file after analysis': True}, result_folder=str(tmp_path)) is True
    assert not csv.exists()
    assert main.ysmr(paths=str(tmp_path / 'v.avi'), settings=dict(st), result_folder=str(tmp_path)) == [(str(tmp_path / 'v.avi'), True)]
