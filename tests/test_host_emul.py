"""The YSMR_HD logic of the CUDA sources (ysmr_b200/csrc/{geometry,label,link}.cuh) compiled for the host by
tests/host_emul and checked against cv2 / scipy / the reference fixtures.  This is what lets the CPU-only suite vouch
for the kernel logic; the -m gpu tests then check the same things through the C-ABI on the device."""
import ctypes
import glob
import os

import cv2
import numpy as np
import pytest
from scipy.ndimage import binary_propagation

from tests.util import GOLDEN, ROW_DT, pack_dets, ptr


def rand_img(rng, h, w, kind):
    if kind == 0:
        a = rng.random((h, w)).astype(np.float32)
        a = cv2.GaussianBlur(a, (0, 0), rng.uniform(1.0, 3.0))
        return ((a > np.quantile(a, rng.uniform(0.5, 0.95))) * 255).astype(np.uint8)
    if kind == 1:
        return ((rng.random((h, w)) < rng.uniform(0.02, 0.6)) * 255).astype(np.uint8)
    if kind == 2:
        img = np.zeros((h, w), np.uint8)
        for _ in range(rng.integers(3, 30)):
            cv2.ellipse(img, (int(rng.integers(0, w)), int(rng.integers(0, h))),
                        (int(rng.integers(1, 9)), int(rng.integers(1, 4))), float(rng.uniform(0, 180)), 0, 360, 255, -1)
        return img
    if kind == 3:
        img = np.zeros((h, w), np.uint8)
        for _ in range(rng.integers(1, 8)):
            c = (int(rng.integers(0, w)), int(rng.integers(0, h))); r = int(rng.integers(3, 30))
            cv2.circle(img, c, r, 255, int(rng.integers(1, 3)))
            if rng.random() < 0.7:
                cv2.circle(img, c, max(1, r // 3), 255, -1 if rng.random() < 0.5 else 1)
            if rng.random() < 0.5:
                cv2.circle(img, c, max(1, r // 6), 255, -1)
        return img
    return np.full((h, w), 255, np.uint8) if rng.random() < 0.5 else np.zeros((h, w), np.uint8)


def test_label_equals_binary_propagation_and_findcontours(emul):
    rng = np.random.default_rng(0)
    n_blobs = 0
    for it in range(600):
        h, w = int(rng.integers(2, 90)), int(rng.integers(2, 140))
        if it % 7 == 0:
            w = 32 * int(rng.integers(1, 4))
        mask = rand_img(rng, h, w, it % 5)
        direct = it % 3 == 0
        markers = None if direct else (mask & ((rng.random((h, w)) < rng.uniform(0.0, 0.2)) * 255).astype(np.uint8))
        out = np.zeros((h, w), np.uint8); fxy = np.zeros((4096, 2), np.int32)
        cnt = np.zeros(1, np.int32); counts = np.zeros(4, np.uint32)
        st = emul.emul_label(ptr(mask), ptr(markers), h, w, 20000, 4096, ptr(out), ptr(fxy), ptr(cnt), ptr(counts))
        assert st == 0
        ref = mask if direct else binary_propagation(markers, mask=mask).astype(np.uint8) * 255   # track_eval.py:211-214
        assert (ref == out).all()
        cs, _ = cv2.findContours(ref, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)               # track_eval.py:273
        first = np.array([c[0, 0] for c in cs], np.int32).reshape(-1, 2)
        assert len(cs) == cnt[0] and (first == fxy[:cnt[0]]).all()
        n_blobs += len(cs)
    assert n_blobs > 5000


def test_label_overflow_flags(emul):
    rng = np.random.default_rng(1)
    mask = ((rng.random((40, 64)) < 0.5) * 255).astype(np.uint8)
    out = np.zeros_like(mask); fxy = np.zeros((4096, 2), np.int32); cnt = np.zeros(1, np.int32); counts = np.zeros(4, np.uint32)
    assert emul.emul_label(ptr(mask), None, 40, 64, 16, 4096, ptr(out), ptr(fxy), ptr(cnt), ptr(counts)) & 1
    assert emul.emul_label(ptr(mask), None, 40, 64, 20000, 3, ptr(out), ptr(fxy), ptr(cnt), ptr(counts)) & 2
    assert cnt[0] == 3


def test_trace_hull_rect_equal_cv2(emul):
    rng = np.random.default_rng(0)
    tot = exact = close = flips = 0
    for it in range(160):
        h, w = int(rng.integers(16, 120)), int(rng.integers(16, 160))
        img = rand_img(rng, h, w, it % 4)
        cs, _ = cv2.findContours(img, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if not cs:
            continue
        first = np.array([c[0, 0] for c in cs], np.int32)
        out = np.zeros((len(cs), 5), np.float32); npts = np.zeros(len(cs), np.int32)
        emul.emul_blob_rects(ptr(img), h, w, ptr(first), len(cs), ptr(out), ptr(npts))
        for i, c in enumerate(cs):
            xy = np.zeros((len(c) + 8, 2), np.int32)
            m = emul.emul_trace(ptr(img), h, w, int(first[i, 0]), int(first[i, 1]), ptr(xy), len(xy))
            assert m == len(c) and (xy[:m] == c[:, 0, :]).all()                  # border following, SIMPLE vertices
            hc = cv2.convexHull(c, clockwise=False)[:, 0, :]
            cc = np.ascontiguousarray(c[:, 0, :].astype(np.int32)); hx = np.zeros((len(c) + 4, 2), np.int32)
            nh = emul.emul_hull(ptr(cc), len(cc), ptr(hx))
            assert nh == len(hc) and (hx[:nh] == hc).all()                       # hull incl. vertex order
            (x, y), (ww, hh), a = cv2.minAreaRect(c)                             # track_eval.py:287
            ref = np.array([x, y, ww, hh, a], np.float32)
            tot += 1
            if (ref == out[i]).all():
                exact += 1
            elif np.abs(ref - out[i]).max() <= 1e-3:
                close += 1
            else:
                flips += 1
    # known residual: exact-area ties decided by rounding; must stay a tiny fraction
    assert tot > 5000 and exact / tot > 0.995 and flips / tot < 1e-3, (tot, exact, close, flips)


def _emul_linker(emul, use_gsff, max_tracks, max_blobs, fps=30.0, max_distance=0.0):
    gz = np.load(os.path.join(GOLDEN, 'gains.npz'))
    gain = np.concatenate([gz['g10'].ravel(), gz['g20'].ravel(), gz['g30'].ravel()]).astype(np.float64)
    n_i = np.array([10, 20, 30], np.int32)
    return emul.emul_link_create(ctypes.c_double(fps), int(use_gsff), 3, ptr(n_i), ptr(gain), max_tracks, max_blobs,
                                 ctypes.c_double(max_distance))


def test_set_order_emulation(emul):
    import random
    random.seed(2)
    for t in range(3000):
        m = random.randint(1, 3000 if t % 20 == 0 else 60)
        used = set(random.sample(range(m), random.randint(0, m)))
        keys = np.array(sorted(set(range(m)) - used), np.int32)
        if len(keys):
            emul.emul_set_order(ptr(keys), len(keys))
        assert list(keys) == list(set(range(m)).difference(used))            # tracker.py:193


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN, 'link_*.npz'))), ids=os.path.basename)
@pytest.mark.parametrize('chunk', [10 ** 9, 37])
def test_link_logic_equals_reference(emul, path, chunk):
    g = np.load(path)
    counts, rows = g['counts'], g['rows']
    blobs = pack_dets(counts, g['dets'])
    h = _emul_linker(emul, bool(g['use_gsff']), 1024, blobs.shape[1])
    out = np.zeros(len(rows) + 100, ROW_DT); nr = np.zeros(1, np.int64); got = []
    for a in range(0, len(counts), chunk):
        b = min(len(counts), a + chunk)
        st = emul.emul_link_chunk(ctypes.c_void_p(h), ptr(np.ascontiguousarray(counts[a:b])), ptr(np.ascontiguousarray(blobs[a:b])),
                                  a, b - a, ptr(out), ctypes.c_longlong(len(out)), ptr(nr))
        assert st == 0
        got.append(out[:nr[0]].copy())
    emul.emul_link_destroy(ctypes.c_void_p(h))
    got = np.concatenate(got)
    assert len(got) == len(rows)
    assert (got['frame'] == rows[:, 0]).all() and (got['track_id'] == rows[:, 1]).all()       # ids bit-exact
    for k, col in (('w', 4), ('h', 5), ('deg', 6)):
        assert (got[k] == rows[:, col].astype(np.float32)).all()
    # positions: bit-identical on every row, coasting tracks included (link.cuh restates NumPy's roundings)
    assert (got['x'] == rows[:, 2]).all() and (got['y'] == rows[:, 3]).all()


def test_link_capacity_flags(emul):
    counts = np.array([5, 5], np.int32)
    blobs = np.zeros((2, 5, 5), np.float32); blobs[:, :, 0] = np.arange(5) * 10
    h = _emul_linker(emul, True, 3, 5)
    out = np.zeros(100, ROW_DT); nr = np.zeros(1, np.int64)
    st = emul.emul_link_chunk(ctypes.c_void_p(h), ptr(counts), ptr(blobs), 0, 2, ptr(out), ctypes.c_longlong(100), ptr(nr))
    assert st & 8 and nr[0] == 6                                             # track overflow, 3 tracks x 2 frames
    emul.emul_link_destroy(ctypes.c_void_p(h))
    h = _emul_linker(emul, True, 16, 5)
    st = emul.emul_link_chunk(ctypes.c_void_p(h), ptr(counts), ptr(blobs), 0, 2, ptr(out), ctypes.c_longlong(7), ptr(nr))
    assert st & 16 and nr[0] == 5                                            # row overflow on the second frame
    emul.emul_link_destroy(ctypes.c_void_p(h))
