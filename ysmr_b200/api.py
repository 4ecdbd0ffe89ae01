"""Python host side of the C-ABI: one :class:`Context` per GPU.

PyTorch is used only for device/pinned buffers and streams; every computation is a call into ``libysmr_b200.so``.
Names follow the reference's domain: frames, blobs (min-area rectangles), tracks, rows (lines of ``_list.csv``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DebugOut, Params, YsmrError

ROW_DTYPE = np.dtype([('frame', '<i4'), ('track_id', '<i4'), ('x', '<f8'), ('y', '<f8'),
                      ('w', '<f4'), ('h', '<f4'), ('deg', '<f4'), ('pad', '<i4')])
assert ROW_DTYPE.itemsize == C.sizeof(_lib.Row) == 40


def horizon_sizes(n_min, n_max, n_f):
    """Filter horizons of the GSFF bank (reference: ysmr/gsff.py:87-109, equation 17)."""
    step = (n_max - n_min) / n_f
    return [int(n_min + step * i) for i in range(1, n_f + 1)]


def lsf_gain(horizon, dt):
    """Least-squares FIR gain (reference: ysmr/gsff.py:112-153, equations 13/14): (L^T L)^-1 L^T with L = H_bar A^-N.

    Computed on the host with the same numpy calls and operation order as the reference so that the float64 gain
    uploaded to the GPU is bit-identical to the one the reference multiplies with.  4 x (2*horizon), C-contiguous."""
    a = np.array([[1, 0, dt, 0], [0, 1, 0, dt], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    c = np.array([[1, 0, 0, 0], [0, 1, 0, 0]])
    h_bar, a_n = c, a
    for _ in range(horizon - 1):
        h_bar = np.concatenate((h_bar, np.dot(c, a_n)), axis=0)
        a_n = np.dot(a_n, a)
    l_bar = np.dot(h_bar, np.linalg.matrix_power(np.linalg.inv(a), horizon))
    return np.ascontiguousarray(np.dot(np.linalg.inv(np.dot(l_bar.T, l_bar)), l_bar.T), dtype=np.float64)


def default_params() -> Params:
    p = Params()
    _lib.load().ysmr_default_params(C.byref(p))
    return p


class Context:
    """Owns one ``ysmr_ctx`` (detection buffers + the persistent linker state) on one GPU."""

    def __init__(self, height, width, channels=1, device=0, *, white_on_dark=True, offset=5, adt=2.0, fps=30.0,
                 use_gsff=True, n_f=3, n_min=0, n_max=30, max_blobs=None, max_tracks=None, max_runs=None,
                 max_batch=None, max_distance=0.0):
        import torch
        if not torch.cuda.is_available():
            raise YsmrError(-2, 'no CUDA device available; ysmr_b200 has no CPU path')
        self.lib = _lib.load()
        self.torch = torch
        p = default_params()
        p.white_on_dark = int(bool(white_on_dark)); p.offset = int(offset); p.adt = float(adt); p.fps = float(fps)
        p.use_gsff = int(bool(use_gsff)); p.n_f = int(n_f); p.n_min = int(n_min)
        p.n_max = int(fps if n_max is None else n_max)          # tracker.py:58-59
        # The horizons themselves come from the reference's expression on the reference's operand types (gsff.py:103-106):
        # with 'maximum horizon size' unset n_max is the FLOAT fps (29.97 fps, four filters: 7, 14, 22, 29 -- int(fps) = 29
        # would give 7, 14, 21, 29)
        self.horizons = horizon_sizes(p.n_min, fps if n_max is None else n_max, p.n_f) if p.use_gsff else []
        for i, n in enumerate(self.horizons[:4]):
            p.reserved[i] = int(n)
        for key, val in (('max_blobs', max_blobs), ('max_tracks', max_tracks), ('max_runs', max_runs),
                         ('max_batch', max_batch)):
            if val is not None:
                setattr(p, key, int(val))
        p.max_distance = float(max_distance)
        self.params = p
        self.height, self.width, self.channels, self.device = int(height), int(width), int(channels), int(device)
        self.tdev = torch.device('cuda', self.device)
        self._h = C.c_void_p()
        rc = self.lib.ysmr_create(C.byref(self._h), self.device, self.height, self.width, self.channels, C.byref(p))
        if rc != 0:
            raise YsmrError(rc, self.lib.ysmr_last_error(None).decode())
        if p.use_gsff:
            for i, n in enumerate(self.horizons):
                g = lsf_gain(n, 1 / p.fps)
                self._check(self.lib.ysmr_set_gsff_gain(self._h, i, n, g.ctypes.data_as(C.c_void_p)))

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise YsmrError(rc, self.lib.ysmr_last_error(self._h).decode())

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self.lib.ysmr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream_ptr(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    def _frames_ok(self, frames):
        t = self.torch
        shape = (self.height, self.width) if self.channels == 1 else (self.height, self.width, 3)
        if frames.dtype != t.uint8 or tuple(frames.shape[1:]) != shape or not frames.is_contiguous():
            raise ValueError(f'frames must be contiguous uint8 of shape (n, {", ".join(map(str, shape))})')
        return int(frames.shape[0]), int(np.prod(shape))

    @property
    def max_batch(self):
        return int(self.params.max_batch)

    @property
    def max_blobs(self):
        return int(self.params.max_blobs)

    # -- detection (track_eval.py:180-303) ----------------------------------------------------------------------
    def detect(self, frames, first_frame=0, debug=False):
        """frames: CUDA uint8 tensor (n, H, W) or (n, H, W, 3), n <= max_batch.
        Returns (blob_count int32[n], blobs float32[n, max_blobs, 5]) CUDA tensors (+ dict of stage images if debug)."""
        t = self.torch
        n, frame_bytes = self._frames_ok(frames)
        assert frames.is_cuda and frames.device == self.tdev
        counts = t.empty(n, dtype=t.int32, device=self.tdev)
        blobs = t.empty((n, self.max_blobs, 5), dtype=t.float32, device=self.tdev)
        dbg_struct, dbg = None, None
        if debug:
            plane = (n, self.height, self.width)
            dbg = {k: t.zeros(plane, dtype=t.uint8, device=self.tdev)
                   for k in ('grey', 'blurred', 'mean', 'mask', 'markers', 'out')}
            dbg['first_xy'] = t.zeros((n, self.max_blobs, 2), dtype=t.int32, device=self.tdev)
            dbg['scalar_thr'] = t.zeros(n, dtype=t.int32, device=self.tdev)
            dbg_struct = DebugOut(*(C.c_void_p(dbg[k].data_ptr()) for k in
                                    ('grey', 'blurred', 'mean', 'mask', 'markers', 'out', 'first_xy', 'scalar_thr')))
        self._check(self.lib.ysmr_detect(self._h, C.c_void_p(frames.data_ptr()), n, frame_bytes, int(first_frame),
                                         C.c_void_p(counts.data_ptr()), C.c_void_p(blobs.data_ptr()),
                                         C.byref(dbg_struct) if dbg_struct is not None else None, self._stream_ptr()))
        return (counts, blobs, dbg) if debug else (counts, blobs)

    # -- linking (tracker.py:93-230, gsff.py) -------------------------------------------------------------------
    def link(self, counts, blobs, first_frame=0, rows_capacity=None):
        """Sequentially links len(counts) frames; returns the rows as a numpy structured array (ROW_DTYPE)."""
        t = self.torch
        n = int(counts.shape[0])
        assert counts.dtype == t.int32 and blobs.dtype == t.float32 and counts.is_cuda and blobs.is_cuda
        assert blobs.shape[1] == self.max_blobs and blobs.is_contiguous() and counts.is_contiguous()
        cap = int(rows_capacity) if rows_capacity is not None else n * int(self.params.max_tracks)
        cap = max(1, min(cap, 1 << 27))
        rows = t.empty(cap * ROW_DTYPE.itemsize, dtype=t.uint8, device=self.tdev)
        n_rows = t.zeros(1, dtype=t.int64, device=self.tdev)
        self._check(self.lib.ysmr_link(self._h, C.c_void_p(counts.data_ptr()), C.c_void_p(blobs.data_ptr()),
                                       int(first_frame), n, C.c_void_p(rows.data_ptr()), cap,
                                       C.c_void_p(n_rows.data_ptr()), self._stream_ptr()))
        self.status()
        k = int(n_rows.item())
        return rows[:k * ROW_DTYPE.itemsize].cpu().numpy().view(ROW_DTYPE).copy()

    def reset(self):
        self._check(self.lib.ysmr_link_reset(self._h))

    def live_tracks(self):
        n, nxt = C.c_int32(), C.c_int32()
        self._check(self.lib.ysmr_link_live_tracks(self._h, C.byref(n), C.byref(nxt)))
        return n.value, nxt.value

    def export_state(self) -> bytes:
        size = C.c_size_t(0)
        self._check(self.lib.ysmr_link_state_export(self._h, None, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        self._check(self.lib.ysmr_link_state_export(self._h, buf, C.byref(size)))
        return buf.raw

    def import_state(self, blob: bytes):
        self._check(self.lib.ysmr_link_state_import(self._h, blob, len(blob)))

    def status(self):
        bits, bad = C.c_int32(), C.c_int32()
        self._check(self.lib.ysmr_status(self._h, self._stream_ptr(), C.byref(bits), C.byref(bad)))

    # -- whole pipeline -----------------------------------------------------------------------------------------
    def track_device(self, frames, first_frame=0, rows_capacity=None, rows_buf=None, return_device=False):
        """detect + link over any number of device-resident frames (chunked, detection overlapping the linker)."""
        t = self.torch
        n, frame_bytes = self._frames_ok(frames)
        cap = int(rows_capacity) if rows_capacity is not None else n * 256
        if rows_buf is None:
            rows_buf = t.empty(max(cap, 1) * ROW_DTYPE.itemsize, dtype=t.uint8, device=self.tdev)
        n_rows = t.zeros(1, dtype=t.int64, device=self.tdev)
        self._check(self.lib.ysmr_track_device(self._h, C.c_void_p(frames.data_ptr()), n, frame_bytes, int(first_frame),
                                               C.c_void_p(rows_buf.data_ptr()), cap, C.c_void_p(n_rows.data_ptr()),
                                               self._stream_ptr()))
        if return_device:
            return rows_buf, n_rows
        self.status()
        k = int(n_rows.item())
        return rows_buf[:k * ROW_DTYPE.itemsize].cpu().numpy().view(ROW_DTYPE).copy()

    def archive_rows(self, enabled=True):
        """Row sink on the device: track_host calls also append their rows to an archive in device memory;
        rows_sorted() returns the whole archive grouped by (track_id, frame) -- the order of the final csv -- sorted on
        the GPU and copied to the host once."""
        self._check(self.lib.ysmr_rows_archive(self._h, 1 if enabled else 0))
        self._archive = bool(enabled)

    def rows_sorted(self):
        k = C.c_int64(0)
        self._check(self.lib.ysmr_rows_sorted(self._h, None, 0, C.byref(k)))
        out = np.empty(max(k.value, 1), ROW_DTYPE)
        self._check(self.lib.ysmr_rows_sorted(self._h, out.ctypes.data_as(C.c_void_p), len(out), C.byref(k)))
        return out[:k.value]

    def track_host(self, frames: np.ndarray, first_frame=0, rows_capacity=None, rows_out=None, copy_rows=True):
        """detect + link over HOST frames (numpy uint8, ideally backed by pinned memory): H2D copies, kernels and the
        D2H copy of the rows all happen inside the one C call.  copy_rows=False (with archive_rows() on): the rows stay
        in the device archive, the number of rows is returned instead."""
        shape = (self.height, self.width) if self.channels == 1 else (self.height, self.width, 3)
        if frames.dtype != np.uint8 or tuple(frames.shape[1:]) != shape or not frames.flags.c_contiguous:
            raise ValueError('frames must be C-contiguous uint8 (n, H, W[, 3])')
        n = int(frames.shape[0])
        cap = int(rows_capacity) if rows_capacity is not None else n * 256
        k = C.c_int64(0)
        if not copy_rows:
            self._check(self.lib.ysmr_track_host(self._h, frames.ctypes.data_as(C.c_void_p), n, int(np.prod(shape)),
                                                 int(first_frame), None, cap, C.byref(k)))
            return int(k.value)
        if rows_out is None:
            rows_out = np.empty(max(cap, 1), ROW_DTYPE)
        self._check(self.lib.ysmr_track_host(self._h, frames.ctypes.data_as(C.c_void_p), n, int(np.prod(shape)),
                                             int(first_frame), rows_out.ctypes.data_as(C.c_void_p), cap, C.byref(k)))
        return rows_out[:k.value]

    OPT_FRONTEND_GEN = 1

    def set_option(self, option, value):
        """Development / measurement switches of the library (include/ysmr_b200.h: YSMR_OPT_*)."""
        self._check(self.lib.ysmr_set_option(self._h, int(option), int(value)))

    def launch_count(self):
        return int(self.lib.ysmr_launch_count(self._h))

    PROF_KINDS = ('frontend', 'label', 'geometry', 'link', 'k1a', 'k1b', 'k1c')

    def set_profiling(self, enabled=True, link_phases=False):
        self._check(self.lib.ysmr_set_profiling(self._h, (1 if enabled else 0) | (2 if link_phases else 0)))

    def link_phase_cycles(self):
        out = (C.c_int64 * 16)()
        self._check(self.lib.ysmr_link_phase_cycles(self._h, out))
        return list(out)

    def get_profile(self):
        """{kind: (total_ms, launches)} since the last call; synchronises the device."""
        ms = (C.c_double * 7)(); n = (C.c_int64 * 7)()
        self._check(self.lib.ysmr_get_profile(self._h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROF_KINDS)}
