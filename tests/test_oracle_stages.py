"""The C restatement of the front-end arithmetic (oracle/c_stages.c) against the pinned third-party implementation
(cv2 4.13) on the reference's call sites (track_eval.py:180, 182, 189-208), and against the golden stage dumps."""
import os

import cv2
import numpy as np
import pytest

from oracle import ref_stages
from tests.util import GOLDEN, c_gauss_mean, gauss_taps, ptr


def test_gauss_taps_match_library_constants():
    # kGaussBits in ysmr_b200/csrc/capi.cu
    bits = [0x3c10612b, 0x3cde5c35, 0x3d855a85, 0x3df92326, 0x3e353f0f, 0x3e4d6105,
            0x3e353f0f, 0x3df92326, 0x3d855a85, 0x3cde5c35, 0x3c10612b]
    assert list(gauss_taps().view(np.uint32)) == bits
    src = open(os.path.join(os.path.dirname(GOLDEN), '..', 'ysmr_b200', 'csrc', 'capi.cu')).read()
    for b in bits:
        assert '0x%08xu' % b in src


def test_grey_bit_exact(cstages):
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (97, 133, 3), dtype=np.uint8)
    out = np.empty((97, 133), np.uint8)
    cstages.ysmr_oracle_grey(ptr(bgr), ptr(out), 97 * 133)
    assert (out == cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)).all()
    g = rng.integers(0, 256, (20, 30), dtype=np.uint8)
    rep = np.ascontiguousarray(np.repeat(g[..., None], 3, -1))
    cstages.ysmr_oracle_grey(ptr(rep), ptr(out), 20 * 30)
    assert (out.ravel()[:600].reshape(20, 30) == g).all()          # identity when B == G == R


@pytest.mark.parametrize('shape', [(16, 16), (37, 53), (120, 164), (99, 157)])
def test_blur3_bit_exact(cstages, shape):
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    out = np.empty(shape, np.uint8)
    cstages.ysmr_oracle_blur3(ptr(src), ptr(out), shape[0], shape[1])
    assert (out == cv2.GaussianBlur(src, (3, 3), 0)).all()


@pytest.mark.parametrize('w', [16, 17, 18, 19, 20, 23, 28, 33, 38, 39, 41, 100, 103, 1228, 1229, 1230, 1231, 2048])
def test_gauss11_float_mean_bit_exact_every_tail(cstages, w):
    rng = np.random.default_rng(w)
    for h in (2, 16, 37):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.GaussianBlur(src.astype(np.float32), (11, 11), 0, 0,
                               borderType=cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED)
        mf, _ = c_gauss_mean(cstages, src)
        assert (mf.view(np.uint32) == ref.view(np.uint32)).all()


@pytest.mark.parametrize('white,offset,adt', [(True, 5, 2.0), (False, 5, 2.0), (True, 4, 1.5), (False, 3, 0.5), (True, 0, 0.0)])
def test_adaptive_masks_bit_exact(cstages, white, offset, adt):
    rng = np.random.default_rng(7)
    src = np.clip(rng.normal(90, 12, (64, 84)), 0, 255).astype(np.uint8)
    blurred = cv2.GaussianBlur(src, (3, 3), 0)
    st = ref_stages.DetectSettings(white, offset, adt)
    ref = ref_stages.threshold_frame(src, st)
    _, mean = c_gauss_mean(cstages, blurred)
    off = st.signed_offset()
    for key, cval in (('mask', off * -1), ('markers', (off + adt) * -1)):
        if ref[key] is None:
            continue
        t = -int(np.ceil(cval)) if white else -int(np.floor(cval))
        got = np.empty_like(src)
        cstages.ysmr_oracle_compare(ptr(blurred), ptr(mean), ptr(got), src.size, t, 0 if white else 1)
        assert (got == ref[key]).all(), key


@pytest.mark.parametrize('name', ['wod', 'dol', 'odd'])
def test_live_cv2_matches_golden_stage_dumps(name):
    """cv2/scipy in THIS process against the dumps made in the build container: detects a different CPU dispatch."""
    g = np.load(os.path.join(GOLDEN, f'stages_{name}.npz'))
    st = ref_stages.DetectSettings(bool(g['white_on_dark']), int(g['offset']), float(g['adt']))
    off = 0
    for i, f in enumerate(g['grey']):
        r = ref_stages.detect_frame(f, st)
        for key in ('blurred', 'mask', 'markers', 'out'):
            assert (r[key] == g[key][i]).all(), (key, i)
        a = ref_stages.rects_to_array(r['rects'])
        assert len(a) == g['counts'][i]
        assert (a == g['rects'][off:off + len(a)]).all()
        off += len(a)
    assert (np.stack([cv2.cvtColor(b, cv2.COLOR_BGR2GRAY) for b in g['bgr']]) == g['bgr_gray']).all()


def test_mean_std_arithmetic_probe():
    """cv2.meanStdDev = (s1*scale, sqrt(max(s2*scale - mean*mean, 0))) with scale = 1/N -- what moving_threshold_kernel does."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        h, w = int(rng.integers(16, 200)), int(rng.integers(16, 300))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        m, s = cv2.meanStdDev(img)
        s1 = float(img.astype(np.int64).sum()); s2 = float((img.astype(np.int64) ** 2).sum())
        scale = 1.0 / float(h * w)
        mean = s1 * scale
        assert mean == m[0, 0]
        assert np.sqrt(max(s2 * scale - mean * mean, 0.0)) == s[0, 0]
