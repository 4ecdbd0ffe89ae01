"""ctypes binding of ``libysmr_b200.so`` (include/ysmr_b200.h).

This is the binding a YSMR maintainer would add next to ``ysmr/track_eval.py``.  The library is
built in-tree by ``ysmr_b200/csrc/Makefile`` (``__graft_entry__.build()``); there is no CPU fallback: if the shared
object is missing or no CUDA device is present, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('YSMR_LIB') or os.path.join(HERE, 'libysmr_b200.so')
CSRC = os.path.join(HERE, 'csrc')

YSMR_OK = 0
ST_RUN_OVERFLOW, ST_BLOB_OVERFLOW, ST_POINT_OVERFLOW, ST_TRACK_OVERFLOW, ST_ROW_OVERFLOW = 1, 2, 4, 8, 16


class YsmrError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'ysmr_b200 error {code}: {message}')
        self.code = code


class Params(C.Structure):
    """struct ysmr_params"""
    _fields_ = [('white_on_dark', C.c_int32), ('offset', C.c_int32), ('adt', C.c_double), ('fps', C.c_double),
                ('use_gsff', C.c_int32), ('n_f', C.c_int32), ('n_min', C.c_int32), ('n_max', C.c_int32),
                ('max_blobs', C.c_int32), ('max_tracks', C.c_int32), ('max_runs', C.c_int32), ('max_batch', C.c_int32),
                ('max_distance', C.c_double), ('reserved', C.c_int32 * 4)]


class Row(C.Structure):
    """struct ysmr_row (40 bytes)"""
    _fields_ = [('frame', C.c_int32), ('track_id', C.c_int32), ('x', C.c_double), ('y', C.c_double),
                ('w', C.c_float), ('h', C.c_float), ('deg', C.c_float), ('pad', C.c_int32)]


class DebugOut(C.Structure):
    """struct ysmr_debug_out"""
    _fields_ = [('d_grey', C.c_void_p), ('d_blurred', C.c_void_p), ('d_mean', C.c_void_p), ('d_mask', C.c_void_p),
                ('d_markers', C.c_void_p), ('d_out', C.c_void_p), ('d_first_xy', C.c_void_p),
                ('d_scalar_thr', C.c_void_p)]


class SelectParams(C.Structure):
    """struct ysmr_select_params"""
    _fields_ = [('area_lo', C.c_double), ('area_hi', C.c_double), ('area_factor', C.c_double), ('q_area', C.c_double),
                ('stop_outliers_above', C.c_double), ('max_empty', C.c_double), ('ratio_min', C.c_double),
                ('ratio_max', C.c_double), ('edge', C.c_double), ('min_len_frames', C.c_int32), ('limit_frames', C.c_int32),
                ('limit_exactly', C.c_int32), ('omit_motility_outliers', C.c_int32), ('max_holes', C.c_int32),
                ('max_recursion', C.c_int32), ('frame_h', C.c_int32), ('frame_w', C.c_int32)]


# indices of the info array of ysmr_select_tracks (include/ysmr_b200.h)
SI_STATUS, SI_ROWS_BEFORE, SI_ROWS_AFTER, SI_TRACKS_BEFORE, SI_TRACKS_AFTER = 0, 1, 2, 3, 4
SI_Q1_AREA, SI_Q3_AREA, SI_Q1_DIST, SI_Q3_DIST, SI_FENCE, SI_OUTLIERS, SI_OUTLIERS_OFF, SI_GOOD_TRACKS, SI_LAUNCHES = range(5, 14)
SELECT_INFO = 16
SEL_OK, SEL_TOO_SHORT_BEFORE, SEL_TOO_SHORT_AFTER, SEL_NO_TRACKS = 0, 1, 2, 3

EXPORTS = {
    # name: (restype, argtypes)
    'ysmr_abi_version': (C.c_int, []),
    'ysmr_last_error': (C.c_char_p, [C.c_void_p]),
    'ysmr_default_params': (None, [C.POINTER(Params)]),
    'ysmr_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params)]),
    'ysmr_destroy': (C.c_int, [C.c_void_p]),
    'ysmr_set_gsff_gain': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'ysmr_detect': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                              C.POINTER(DebugOut), C.c_void_p]),
    'ysmr_link': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                            C.c_void_p]),
    'ysmr_link_append': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p]),
    'ysmr_link_reset': (C.c_int, [C.c_void_p]),
    'ysmr_link_state_export': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    'ysmr_link_state_import': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    'ysmr_link_live_tracks': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'ysmr_status': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'ysmr_rows_archive': (C.c_int, [C.c_void_p, C.c_int]),
    'ysmr_rows_sorted': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    'ysmr_track_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                  C.POINTER(C.c_int64)]),
    'ysmr_track_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
    'ysmr_set_option': (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    'ysmr_launch_count': (C.c_int64, [C.c_void_p]),
    'ysmr_set_profiling': (C.c_int, [C.c_void_p, C.c_int]),
    'ysmr_get_profile': (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    'ysmr_link_phase_cycles': (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    'ysmr_select_tracks': (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(SelectParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ysmr_select_last_error': (C.c_char_p, []),
    'ysmr_track_statistics': (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int32, C.c_void_p]),
    'ysmr_statistics_last_error': (C.c_char_p, []),
}

_lib = None


def build(verbose=False):
    """Compile libysmr_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ['make', '-C', CSRC, '-j', str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError('building libysmr_b200.so failed')
    return LIB_PATH


def load():
    """Load the shared library (no compute happens here, so this also works on a box without a GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FileNotFoundError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` or '
            f'`make -C {CSRC}`.  ysmr_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
