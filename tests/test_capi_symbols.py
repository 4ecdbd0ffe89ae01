"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/ysmr_b200.h declares."""
import ctypes
import os
import re

from tests.util import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'ysmr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ysmr_[a-z_0-9]+)\s*\(', text)))


def test_library_builds_loads_and_exports_the_header():
    from ysmr_b200 import _lib
    _lib.build()
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert lib.ysmr_abi_version() == 5


def test_struct_layouts_match_header():
    from ysmr_b200 import _lib
    from ysmr_b200.api import ROW_DTYPE
    assert ctypes.sizeof(_lib.Row) == 40 == ROW_DTYPE.itemsize
    assert ctypes.sizeof(_lib.Params) == 4 * 2 + 8 * 2 + 4 * 8 + 8 + 4 * 4
    lib = _lib.load()
    p = _lib.Params()
    lib.ysmr_default_params(ctypes.byref(p))
    # reference defaults, helper_file.py:160-282
    assert (p.white_on_dark, p.offset, p.adt, p.fps, p.use_gsff, p.n_f, p.n_min, p.n_max) == (1, 5, 2.0, 30.0, 1, 3, 0, 30)
    assert p.max_distance == 0.0        # the reference has no gate (tracker.py:171-177)


def test_create_fails_loudly_without_gpu_or_with_bad_geometry():
    import torch
    from ysmr_b200 import _lib
    lib = _lib.load()
    p = _lib.Params(); lib.ysmr_default_params(ctypes.byref(p))
    h = ctypes.c_void_p()
    assert lib.ysmr_create(ctypes.byref(h), 0, 8, 8, 1, ctypes.byref(p)) == -1          # too small
    assert b'16' in lib.ysmr_last_error(None)
    assert lib.ysmr_create(ctypes.byref(h), 0, 64, 64, 2, ctypes.byref(p)) == -1         # channels
    if not torch.cuda.is_available():
        assert lib.ysmr_create(ctypes.byref(h), 0, 64, 64, 1, ctypes.byref(p)) == -2     # no CUDA device: no fallback
        assert not h.value
