"""Shared helpers of the test-suite (test infrastructure; may use oracle/)."""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
P = ctypes.c_void_p


def ptr(a):
    return None if a is None else a.ctypes.data_as(P)


def _stale(out, deps):
    return not os.path.isfile(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps)


def build_cstages():
    src = os.path.join(ROOT, 'oracle', 'c_stages.c')
    out = os.path.join(ROOT, 'oracle', '_build', 'c_stages.so')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if _stale(out, [src]):
        subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', out, src, '-lm'], check=True)
    return out


def build_emul():
    src = os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp')
    deps = [src] + glob.glob(os.path.join(ROOT, 'ysmr_b200', 'csrc', '*.cuh'))
    out = os.path.join(ROOT, 'tests', 'host_emul', '_build', 'emul.so')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if _stale(out, deps):
        subprocess.run(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC', '-o', out, src], check=True)
    return out


def load_cstages():
    return ctypes.CDLL(build_cstages())


def load_emul():
    lib = ctypes.CDLL(build_emul())
    lib.emul_link_create.restype = ctypes.c_void_p
    lib.emul_blas_row_dot.restype = ctypes.c_double
    return lib


def gauss_taps():
    import cv2
    return cv2.getGaussianKernel(11, 0, cv2.CV_32F).reshape(-1).copy()


def c_gauss_mean(lib, src):
    h, w = src.shape
    mf = np.empty((h, w), np.float32); mu = np.empty((h, w), np.uint8)
    k = gauss_taps()
    lib.ysmr_oracle_gauss11_mean(ptr(np.ascontiguousarray(src)), ptr(k), ptr(mf), ptr(mu), h, w)
    return mf, mu


ROW_DT = np.dtype([('frame', '<i4'), ('track_id', '<i4'), ('x', '<f8'), ('y', '<f8'),
                   ('w', '<f4'), ('h', '<f4'), ('deg', '<f4'), ('pad', '<i4')])


def pack_dets(counts, dets, max_blobs=None):
    """ragged (counts, flat dets) -> dense [n_frames, max_blobs, 5] float32"""
    mb = int(max(counts.max() if len(counts) else 1, 1)) if max_blobs is None else max_blobs
    blobs = np.zeros((len(counts), mb, 5), np.float32)
    off = 0
    for t, c in enumerate(counts):
        blobs[t, :c] = dets[off:off + c]
        off += c
    return blobs


def oracle_rows(frames, settings, fps=30.0, use_gsff=True):
    """Full oracle pipeline (cv2/scipy call sites + tracker port) over frames; rows in emission order
    [(frame, id, x, y, w, h, deg)] plus per-frame rect arrays."""
    from oracle import ref_stages
    from oracle.tracker_port import LinkerPort
    lp = LinkerPort(max_disappeared=fps, fps=fps, use_gsff=use_gsff)
    rows, per_frame = [], []
    for t, f in enumerate(frames):
        r = ref_stages.detect_frame(f, settings)
        per_frame.append(ref_stages.rects_to_array(r['rects']))
        for (i, xy, info) in lp.update(r['rects']):
            rows.append((t, i, xy[0], xy[1], info[0], info[1], info[2]))
    return np.array(rows, np.float64).reshape(-1, 7), per_frame


def coasting_age(info, track_ids, window=32):
    """For each row (emission order): the longest run of consecutive UNMATCHED frames its track had within the last
    `window` frames.  info = (n, 3) array of (w, h, deg).  A track is unmatched in a frame when its info is zeroed
    (tracker.py:101, 205) or -- in the "more detections than tracks" branch, which neither ages nor zeroes
    (tracker.py:215-217) -- when it simply repeats the previous frame's info.

    Why tests need this: while a track is unmatched the reference feeds the GSFF prediction back as the next
    measurement (tracker.py:219-227); that loop amplifies last-bit differences (BLAS summation order, exp) by ~6x per
    frame, and the affected measurements stay in the 31-frame filter history after the track is matched again."""
    info = np.asarray(info)
    age = np.zeros(len(track_ids), np.int32)
    cur, hist, prev = {}, {}, {}
    for i, tid in enumerate(track_ids):
        row = tuple(info[i])
        unmatched = (row == (0.0, 0.0, 0.0)) or (prev.get(tid) == row)
        prev[tid] = row
        a = cur.get(tid, 0) + 1 if unmatched else 0
        cur[tid] = a
        h = hist.setdefault(tid, [])
        h.append(a)
        if len(h) > window:
            del h[0]
        age[i] = max(h)
    return age
