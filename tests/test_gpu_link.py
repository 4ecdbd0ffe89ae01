"""-m gpu: the persistent linker kernel through the C-ABI against fixtures from the reference CentroidTracker /
GaussianSumFIR (tracker.py:93-230, gsff.py:204-347) and against the oracle port on fresh random sequences."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from tests.util import GOLDEN, pack_dets  # noqa: E402


def _run(counts, blobs, use_gsff=True, chunk=10 ** 9, fps=30.0, max_tracks=2048, **kw):
    from ysmr_b200.api import Context
    ctx = Context(64, 64, 1, 0, fps=fps, use_gsff=use_gsff, max_blobs=blobs.shape[1], max_tracks=max_tracks, max_batch=4, **kw)
    c = torch.from_numpy(counts).cuda(); b = torch.from_numpy(blobs).cuda()
    got = []
    for a in range(0, len(counts), chunk):
        e = min(len(counts), a + chunk)
        got.append(ctx.link(c[a:e].contiguous(), b[a:e].contiguous(), a, rows_capacity=(e - a) * max_tracks))
    live = ctx.live_tracks()
    ctx.close()
    return np.concatenate(got), live


def _check(got, rows):
    assert len(got) == len(rows)
    assert (got['frame'] == rows[:, 0]).all() and (got['track_id'] == rows[:, 1]).all()      # ids bit-exact
    for k, col in (('w', 4), ('h', 5), ('deg', 6)):
        assert (got[k] == rows[:, col].astype(np.float32)).all()
    # GSFF positions: the device restates the reference's NumPy roundings (csrc/link.cuh), so EVERY row -- coasting
    # tracks included, no exemption -- carries the reference's float64 bits (north-star bar: 1e-5 relative)
    assert (got['x'] == rows[:, 2]).all() and (got['y'] == rows[:, 3]).all(), \
        np.maximum(np.abs(got['x'] - rows[:, 2]), np.abs(got['y'] - rows[:, 3])).max()


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN, 'link_*.npz'))), ids=os.path.basename)
@pytest.mark.parametrize('chunk', [10 ** 9, 37])
def test_golden_sequences(path, chunk):
    g = np.load(path)
    blobs = pack_dets(g['counts'], g['dets'])
    got, _ = _run(g['counts'], blobs, bool(g['use_gsff']), chunk)
    _check(got, g['rows'])


def test_fresh_random_sequences_vs_oracle_port():
    from oracle.make_golden import random_detection_sequence
    from oracle.tracker_port import LinkerPort
    for seed, kw in ((11, dict(n_frames=200, n_cells=25)), (12, dict(n_frames=80, n_cells=600, width=1228., height=922.)),
                     (13, dict(n_frames=150, n_cells=5, p_empty=0.3, p_miss=0.4))):
        rng = np.random.default_rng(seed)
        seq = random_detection_sequence(rng, **kw)
        counts = np.array([len(s) for s in seq], np.int32)
        blobs = pack_dets(counts, np.concatenate(seq))
        lp = LinkerPort(max_disappeared=30.0, fps=30.0)
        rows = []
        for t, rec in enumerate(seq):
            rects = [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]), float(r[4]))) for r in rec]
            rows += [(t, i, xy[0], xy[1], info[0], info[1], info[2]) for (i, xy, info) in lp.update(rects)]
        got, live = _run(counts, blobs)
        _check(got, np.array(rows, np.float64))
        assert live == (len(lp.ids), lp.next_id)


def test_state_export_import_resumes_identically():
    from ysmr_b200.api import Context
    g = np.load(os.path.join(GOLDEN, 'link_b_gsff.npz'))
    blobs = pack_dets(g['counts'], g['dets'])
    c = torch.from_numpy(g['counts']).cuda(); b = torch.from_numpy(blobs).cuda()
    a = Context(64, 64, 1, 0, max_blobs=blobs.shape[1], max_tracks=512, max_batch=4)
    r1 = a.link(c[:100].contiguous(), b[:100].contiguous(), 0, 100 * 512)
    blob = a.export_state()
    r2 = a.link(c[100:].contiguous(), b[100:].contiguous(), 100, 200 * 512)
    z = Context(64, 64, 1, 0, max_blobs=blobs.shape[1], max_tracks=512, max_batch=4)
    z.import_state(blob)
    r3 = z.link(c[100:].contiguous(), b[100:].contiguous(), 100, 200 * 512)
    assert r2.tobytes() == r3.tobytes()
    _check(np.concatenate([r1, r2]), g['rows'])
    a.reset()
    r4 = a.link(c[:100].contiguous(), b[:100].contiguous(), 0, 100 * 512)
    assert r4.tobytes() == r1.tobytes()
    a.close(); z.close()


def test_track_and_row_overflow_are_loud():
    from ysmr_b200._lib import YsmrError
    from ysmr_b200.api import Context
    counts = np.array([5, 5], np.int32)
    blobs = np.zeros((2, 5, 5), np.float32); blobs[:, :, 0] = np.arange(5) * 10
    ctx = Context(64, 64, 1, 0, max_blobs=5, max_tracks=3, max_batch=4)
    with pytest.raises(YsmrError):
        ctx.link(torch.from_numpy(counts).cuda(), torch.from_numpy(blobs).cuda(), 0, 100)
    ctx.close()
    ctx = Context(64, 64, 1, 0, max_blobs=5, max_tracks=16, max_batch=4)
    with pytest.raises(YsmrError):
        ctx.link(torch.from_numpy(counts).cuda(), torch.from_numpy(blobs).cuda(), 0, 7)
    ctx.close()


@pytest.mark.parametrize('mode', ['general,cta', 'general,grid', 'grid'])
def test_general_kernels_single_cta_and_cooperative_grid(monkeypatch, mode):
    """The general path (any number of tracks / detections) as one CTA and as a cooperative grid, with and without the lane
    fast path in front of it: same rows as the reference on the golden sequences and on a fresh dense one."""
    from oracle.make_golden import random_detection_sequence
    from oracle.tracker_port import LinkerPort
    monkeypatch.setenv('YSMR_LINK', mode)
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'link_*_gsff.npz'))):
        g = np.load(path)
        got, _ = _run(g['counts'], pack_dets(g['counts'], g['dets']), True, 53)
        _check(got, g['rows'])
    rng = np.random.default_rng(21)
    seq = random_detection_sequence(rng, n_frames=70, n_cells=900, width=1228., height=922., p_miss=0.05)
    counts = np.array([len(s) for s in seq], np.int32)
    lp = LinkerPort(max_disappeared=30.0, fps=30.0)
    rows = []
    for t, rec in enumerate(seq):
        rects = [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]), float(r[4]))) for r in rec]
        rows += [(t, i, xy[0], xy[1], info[0], info[1], info[2]) for (i, xy, info) in lp.update(rects)]
    got, live = _run(counts, pack_dets(counts, np.concatenate(seq)), True, 16)
    _check(got, np.array(rows, np.float64))
    assert live == (len(lp.ids), lp.next_id)
