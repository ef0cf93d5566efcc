"""CPU test: the C-ABI library builds, loads and exports every symbol include/*.h declares;
without a GPU the compute entry points must refuse (no CPU fallback)."""
import ctypes
import os
import re

from seissol_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "seissol_b200.h")).read()
    return sorted(set(re.findall(r"\b(ssb_[a-z_]+)\s*\(", text)))


def test_library_exports_all_declared_symbols(built_library):
    names = declared_symbols()
    assert set(names) == set(api.SYMBOLS)
    for n in names:
        assert hasattr(built_library, n), n


def test_error_strings_and_version(built_library):
    assert built_library.ssb_error_string(0) == b"success"
    assert b"invalid" in built_library.ssb_error_string(-1)
    assert b"sm_100a" in built_library.ssb_version()


def test_create_rejects_bad_config_or_missing_device(built_library):
    h = ctypes.c_void_p()
    # order 9 is not compiled in
    assert built_library.ssb_create(ctypes.byref(h), 9, 4, 0, 0) == -2
    assert built_library.ssb_create(ctypes.byref(h), 6, 2, 0, 0) == -2
    import torch
    if not torch.cuda.is_available():
        # no CPU fallback: creating a context without a device fails loudly
        assert built_library.ssb_create(ctypes.byref(h), 6, 4, 0, 0) == -3
        try:
            api.Context(6, 4)
            raised = False
        except api.SsbError:
            raised = True
        assert raised
