"""-m gpu: detection through the C-ABI on the device against the oracle (cv2/scipy call sites of
track_eval.py:180-303) on the same seeded bytes, and against the golden stage dumps."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import ref_stages  # noqa: E402
from tests.util import GOLDEN  # noqa: E402
from ysmr_b200.synth import SceneConfig, make_scene, render_frames, to_bgr  # noqa: E402


def _ctx(h, w, channels=1, **kw):
    from ysmr_b200.api import Context
    kw.setdefault('max_batch', 16); kw.setdefault('max_blobs', 2048); kw.setdefault('max_tracks', 4096)
    return Context(h, w, channels, 0, **kw)


def _compare_stages(frames_np, st, channels=1, **ctx_kw):
    """Runs frames through ysmr_detect with debug dumps and the oracle; returns mismatch report."""
    n = len(frames_np)
    h, w = frames_np.shape[1:3]
    ctx = _ctx(h, w, channels, white_on_dark=st.white_on_dark, offset=st.offset, adt=st.adt, fps=st.fps, **ctx_kw)
    fr = torch.from_numpy(frames_np).cuda()
    counts, blobs, dbg = ctx.detect(fr, 0, debug=True)
    ctx.status()
    counts = counts.cpu().numpy(); blobs = blobs.cpu().numpy()
    dbg = {k: v.cpu().numpy() for k, v in dbg.items()}
    rep = {'rect_exact': 0, 'rect_close': 0, 'rect_flip': 0, 'rects': 0}
    for i in range(n):
        r = ref_stages.detect_frame(frames_np[i], st)
        assert (dbg['grey'][i] == r['gray']).all(), f'grey frame {i}'
        assert (dbg['blurred'][i] == r['blurred']).all(), f'blurred frame {i}'
        assert (dbg['mask'][i] == r['mask']).all(), f'mask frame {i}: {(dbg["mask"][i] != r["mask"]).sum()} px'
        if r['markers'] is not None:
            assert (dbg['markers'][i] == r['markers']).all(), f'markers frame {i}'
        assert (dbg['out'][i] == r['out']).all(), f'propagated image frame {i}'
        ref = ref_stages.rects_to_array(r['rects'])
        assert counts[i] == len(ref), f'blob count frame {i}: {counts[i]} vs {len(ref)}'
        fp = ref_stages.first_pixels(r['contours'])
        assert (dbg['first_xy'][i, :len(ref)] == fp).all(), f'contour order / first pixels frame {i}'
        got = blobs[i, :len(ref)]
        for a, b in zip(got, ref):
            rep['rects'] += 1
            if (a == b).all():
                rep['rect_exact'] += 1
            elif np.abs(a - b).max() <= 1e-3:          # 1e-3 px; 1e-3 rad = 0.057 deg, we hold 1e-3 deg
                rep['rect_close'] += 1
            else:
                rep['rect_flip'] += 1
    ctx.close()
    return rep


@pytest.mark.parametrize('name', ['wod', 'dol', 'odd'])
def test_golden_stage_dumps(name):
    g = np.load(os.path.join(GOLDEN, f'stages_{name}.npz'))
    st = ref_stages.DetectSettings(bool(g['white_on_dark']), int(g['offset']), float(g['adt']))
    rep = _compare_stages(g['grey'], st)
    assert rep['rect_flip'] == 0


def test_bgr_luma_bit_exact():
    g = np.load(os.path.join(GOLDEN, 'stages_wod.npz'))
    st = ref_stages.DetectSettings(True, 5, 2.0)
    rep = _compare_stages(np.ascontiguousarray(g['bgr']), st, channels=3)     # true colour: exercises cvtColor's Q15 luma
    assert rep['rect_flip'] <= 1


@pytest.mark.parametrize('white,offset,adt,w', [(True, 5, 2.0, 1228), (False, 5, 2.0, 1228), (True, 5, 0.0, 1228),
                                                (True, 5, 2.0, 1230), (True, 4, 1.5, 1225), (False, 3, 1.0, 2048)])
def test_full_width_frames_all_modes(white, offset, adt, w):
    kw = dict(width=w, height=300 if w != 2048 else 256, n_frames=3, n_cells=40, seed=21, margin=20.0)
    if not white:
        kw.update(background=160.0, intensity=80.0, noise_sigma=1.5, semi_major=2.5, semi_minor=2.5)
    cfg = SceneConfig(**kw)
    grey = render_frames(make_scene(cfg))
    rep = _compare_stages(grey, ref_stages.DetectSettings(white, offset, adt))
    assert rep['rects'] > 50 and rep['rect_flip'] <= 1, rep


@pytest.mark.parametrize('w', [333, 334, 336, 260])
def test_bgr_true_colour_any_width(w):
    """cvtColor + blur pre-pass on true-colour frames of widths that exercise the generic (unaligned) loads, the half-group
    lane (w % 8 == 4) and the scalar-tail lanes of the Gaussian."""
    rng = np.random.default_rng(w)
    cfg = SceneConfig(width=w, height=120, n_frames=3, n_cells=10, seed=w, margin=15.0)
    grey = render_frames(make_scene(cfg)).astype(np.int16)
    bgr = np.stack([np.clip(grey + rng.integers(-20, 21, grey.shape), 0, 255) for _ in range(3)], -1).astype(np.uint8)
    rep = _compare_stages(np.ascontiguousarray(bgr), ref_stages.DetectSettings(True, 5, 2.0), channels=3)
    assert rep['rect_flip'] <= 1, rep


def test_cfg1_full_frames_bgr():
    cfg = SceneConfig(n_frames=4)
    grey = render_frames(make_scene(cfg))
    rep = _compare_stages(to_bgr(grey), ref_stages.DetectSettings(), channels=3)
    assert rep['rects'] >= 190 and rep['rect_flip'] == 0, rep


def test_noise_and_shapes_stress():
    """Heavy noise (thousands of specks, nested rings, border-touching blobs) at the capacities' scale."""
    rng = np.random.default_rng(5)
    import cv2
    frames = []
    for i in range(4):
        img = np.clip(rng.normal(70, 9, (256, 320)), 0, 255).astype(np.uint8)
        for _ in range(12):
            c = (int(rng.integers(0, 320)), int(rng.integers(0, 256))); r = int(rng.integers(5, 40))
            cv2.circle(img, c, r, 200, int(rng.integers(1, 4)))
            if rng.random() < 0.6:
                cv2.circle(img, c, max(1, r // 3), 210, -1)
        frames.append(img)
    rep = _compare_stages(np.stack(frames), ref_stages.DetectSettings(True, 5, 2.0), max_runs=60000, max_blobs=8192)
    assert rep['rects'] > 500 and rep['rect_flip'] <= 2, rep


def test_mean_std_mode_matches_reference_threshold_sequence():
    cfg = SceneConfig(width=200, height=160, n_frames=40, n_cells=8, seed=6, margin=30.0)
    grey = render_frames(make_scene(cfg))
    st = ref_stages.DetectSettings(True, 25, -1.0, fps=5.0)        # window = 26 frames, so it slides within 40 frames
    ctx = _ctx(160, 200, 1, white_on_dark=True, offset=25, adt=-1.0, fps=5.0)
    fr = torch.from_numpy(grey).cuda()
    thr_ref, outs = [], []
    for f in grey:
        r = ref_stages.detect_frame(f, st)
        thr_ref.append(r['scalar_threshold']); outs.append(r)
    got_thr = []
    for a in range(0, 40, 16):                                     # chunks: history must carry over
        counts, blobs, dbg = ctx.detect(fr[a:a + 16], a, debug=True)
        got_thr += dbg['scalar_thr'].cpu().tolist()
        m = dbg['mask'].cpu().numpy()
        for i in range(m.shape[0]):
            assert (m[i] == outs[a + i]['mask']).all()
            assert counts[i].item() == len(outs[a + i]['rects'])
    assert got_thr == thr_ref
    ctx.close()


def test_capacity_overflow_is_loud():
    from ysmr_b200._lib import YsmrError
    rng = np.random.default_rng(0)
    img = (rng.random((2, 128, 160)) * 255).astype(np.uint8)       # pure noise: thousands of runs
    ctx = _ctx(128, 160, 1, max_runs=64, max_blobs=16)
    ctx.detect(torch.from_numpy(img).cuda(), 0)
    with pytest.raises(YsmrError) as e:
        ctx.status()
    assert e.value.code == -3
    ctx.close()
