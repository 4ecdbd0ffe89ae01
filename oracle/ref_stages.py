"""ORACLE (test infrastructure only -- never imported by the product path ``ysmr_b200/``).

Stage-level oracle for the per-frame *detection* half of YSMR's hot loop.  The reference has no native code:
its arithmetic lives in un-vendored third-party wheels (setup.py:47-54 pins only lower
bounds: opencv-contrib-python>=3.4.1, scipy>=1.3.0).  The wheels installed in this image -- OpenCV 4.13.0
(opencv-python-headless 4.13.0.92) and SciPy 1.18.1 -- are therefore the pinned implementation, and this
module replays the reference's own *call sites* against them, one call per reference line:

    track_eval.py:127-132  polarity + in-place sign flip of the offset
    track_eval.py:180      cv2.cvtColor(frame, COLOR_BGR2GRAY)
    track_eval.py:182      cv2.GaussianBlur(gray, (3, 3), 0)
    track_eval.py:189-197  cv2.adaptiveThreshold(... GAUSSIAN_C, type, 11, -offset)            (mask)
    track_eval.py:200-208  cv2.adaptiveThreshold(... GAUSSIAN_C, type, 11, -(offset + adt))    (markers)
    track_eval.py:211-214  scipy.ndimage.binary_propagation(markers, mask=thresh) * 255
    track_eval.py:219-253  mean/std moving-average threshold (adaptive double threshold < 0)
    track_eval.py:273-283  cv2.findContours(thresh, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
    track_eval.py:286-303  cv2.minAreaRect(contour) + helper_file.reshape_result (helper_file.py:1336-1347)

Parity pin: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself).  The
pin used instead is the reference executed in the build container: ``oracle/make_golden.py`` imports
``/root/reference/ysmr`` and stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
this module against them.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import cv2
import numpy as np
from scipy.ndimage import binary_propagation

BLOCK = 11  # kernel_for_thresholding, track_eval.py:187


@dataclass
class DetectSettings:
    """The hot-path keys of tracking.ini with the reference defaults
    (helper_file.py:160-282)."""
    white_on_dark: bool = True           # 'white bacteria on dark background'
    offset: int = 5                      # 'threshold offset for detection'
    adt: float = 2.0                     # 'adaptive double threshold'
    fps: float = 30.0
    threshold_list: list = field(default_factory=list)  # state of the mean/std mode (track_eval.py:120)

    def signed_offset(self) -> int:
        # track_eval.py:127-132: dark-on-light negates the offset once, before the loop
        return self.offset if self.white_on_dark else -self.offset

    def threshold_type(self) -> int:
        return cv2.THRESH_BINARY if self.white_on_dark else cv2.THRESH_BINARY_INV


def grey_of(frame: np.ndarray) -> np.ndarray:
    """track_eval.py:180.  A 2-D frame is taken as the already-grey plane (B=G=R makes BGR2GRAY the identity)."""
    if frame.ndim == 2:
        return frame
    return cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)


def threshold_frame(frame: np.ndarray, st: DetectSettings) -> dict:
    """track_eval.py:180-253 for one frame.  Returns every intermediate so each layer can be compared."""
    gray = grey_of(frame)
    blurred = cv2.GaussianBlur(gray, (3, 3), 0)
    off = st.signed_offset()
    ttype = st.threshold_type()
    res = {'gray': gray, 'blurred': blurred, 'markers': None}
    if st.adt >= 0:
        thresh = cv2.adaptiveThreshold(blurred, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, ttype, BLOCK, off * -1)
        res['mask'] = thresh
        if st.adt > 0:
            markers = cv2.adaptiveThreshold(blurred, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, ttype, BLOCK,
                                            (off + st.adt) * -1)
            res['markers'] = markers
            thresh = binary_propagation(markers, mask=thresh).astype(np.uint8) * 255
    else:
        mean, stddev = cv2.meanStdDev(gray)
        if st.white_on_dark:
            cur = mean + stddev + off
        else:
            cur = mean - stddev - off
        st.threshold_list.append(cur)
        curr_threshold = int((sum(st.threshold_list) / len(st.threshold_list)).item())
        if len(st.threshold_list) > st.fps * 5:
            del st.threshold_list[0]
        res['scalar_threshold'] = curr_threshold
        thresh = cv2.threshold(blurred, curr_threshold, 255, ttype)[1]
        res['mask'] = thresh
    res['out'] = thresh
    return res


def rects_of(thresh: np.ndarray):
    """track_eval.py:273-303: external contours in OpenCV's order, one min-area rectangle per contour,
    reshaped to ((x, y), (w, h, deg)) like helper_file.reshape_result."""
    contours = cv2.findContours(thresh, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    contours = contours[1] if len(contours) == 3 else contours[0]
    rects = []
    for c in contours:
        (x, y), (w, h), deg = cv2.minAreaRect(c)
        rects.append(((x, y), (w, h, deg)))
    return contours, rects


def detect_frame(frame: np.ndarray, st: DetectSettings) -> dict:
    res = threshold_frame(frame, st)
    contours, rects = rects_of(res['out'])
    res['contours'] = contours
    res['rects'] = rects
    return res


def rects_to_array(rects) -> np.ndarray:
    """(n, 5) float32 [cx, cy, w, h, deg] -- the blob record layout of the C-ABI (include/ysmr_b200.h)."""
    a = np.zeros((len(rects), 5), np.float32)
    for i, ((x, y), (w, h, d)) in enumerate(rects):
        a[i] = (x, y, w, h, d)
    return a


def first_pixels(contours) -> np.ndarray:
    """Raster-first pixel of each contour = its first point (cv2 starts every outer border there)."""
    return np.array([c[0, 0] for c in contours], np.int32).reshape(-1, 2)
