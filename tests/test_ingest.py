"""CPU: the drop-in's ingest (ysmr_b200/ingest.py) delivers exactly the frames cap.read() delivers
(track_eval.py:65,159), in order, whether one reader decodes sequentially or several decode disjoint chunks of an intra-only
file in parallel; grey sources come as one verified plane, a colour frame is reported."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')
pytest.importorskip('torch')

from ysmr_b200 import ingest  # noqa: E402


def _write(path, frames_bgr, fourcc='FFV1', fps=30.0):
    h, w = frames_bgr.shape[1:3]
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), fps, (w, h), True)
    assert vw.isOpened()
    for f in frames_bgr:
        vw.write(f)
    vw.release()


def _sequential(path):
    cap = cv2.VideoCapture(str(path))
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    out = []
    while True:
        ret, f = cap.read()
        if not ret:
            break
        out.append(f)
    cap.release()
    return n, np.stack(out)


def _collect(reader):
    got, lasts = [], []
    for buf, idx, n, last in reader:
        got.append(buf.copy()); lasts.append(last)
        reader.release(idx)
    reader.close()
    return np.concatenate(got) if got else np.empty((0,)), lasts


@pytest.mark.parametrize('n_frames,chunk,readers', [(37, 8, 1), (37, 8, 3), (40, 8, 4), (5, 8, 4), (64, 16, 2)])
def test_chunks_equal_cap_read(tmp_path, n_frames, chunk, readers):
    rng = np.random.default_rng(n_frames + readers)
    grey = rng.integers(0, 256, (n_frames, 48, 64), dtype=np.uint8)
    path = tmp_path / 'v.avi'
    _write(path, np.repeat(grey[..., None], 3, axis=-1))
    count, ref = _sequential(path)
    assert len(ref) == n_frames and (ref[..., 0] == grey).all()           # FFV1 round-trips bit-exactly
    r = ingest.ChunkReader(str(path), count, 48, 64, 3, chunk, n_readers=readers)
    assert r.n_readers == (readers if readers > 1 else 1)
    got, lasts = _collect(r)
    assert got.shape == ref.shape and (got == ref).all()
    assert lasts[-1] and not any(lasts[:-1])
    # grey source: one plane, every frame verified
    r1 = ingest.ChunkReader(str(path), count, 48, 64, 1, chunk, n_readers=readers)
    got1, _ = _collect(r1)
    assert got1.shape == grey.shape and (got1 == grey).all()


def test_colour_frame_in_grey_mode_is_reported(tmp_path):
    rng = np.random.default_rng(1)
    grey = rng.integers(0, 256, (20, 32, 32), dtype=np.uint8)
    bgr = np.repeat(grey[..., None], 3, axis=-1)
    bgr[13, 5, 7, 2] ^= 0x40                                              # one red-channel pixel of frame 13
    path = tmp_path / 'c.avi'
    _write(path, bgr)
    count, ref = _sequential(path)
    assert not ingest.is_grey_frame(ref[13]) and ingest.is_grey_frame(ref[12])
    r = ingest.ChunkReader(str(path), count, 32, 32, 1, 8, n_readers=2)
    with pytest.raises(ingest.ColourFrame):
        _collect(r)
    r.close()


def test_inter_frame_codec_gets_a_single_reader(tmp_path):
    grey = np.zeros((12, 32, 32), np.uint8)
    path = tmp_path / 'm.avi'
    _write(path, np.repeat(grey[..., None], 3, axis=-1), fourcc='mp4v')
    cap = cv2.VideoCapture(str(path))
    if not cap.isOpened() or int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 0:
        pytest.skip('mp4v not available in this OpenCV build')
    count = int(cap.get(cv2.CAP_PROP_FRAME_COUNT)); cap.release()
    r = ingest.ChunkReader(str(path), count, 32, 32, 3, 4, n_readers=4)
    assert r.n_readers == 1
    got, _ = _collect(r)
    assert len(got) == 12
