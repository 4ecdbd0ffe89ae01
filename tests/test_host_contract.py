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
