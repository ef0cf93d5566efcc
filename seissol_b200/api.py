"""ctypes binding of the C ABI (include/seissol_b200.h -> seissol_b200/lib/libseissol_b200.so).

PyTorch is used only as the owner of device memory and streams; every compute call
goes through the C ABI into the hand-written sm_100a kernels.  There is no CPU
fallback: a missing library or a missing CUDA device raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libseissol_b200.so")

SYMBOLS = [
    "ssb_create", "ssb_destroy", "ssb_set_layout", "ssb_layout_reference", "ssb_query",
    "ssb_ader", "ssb_local", "ssb_ader_local", "ssb_taylor", "ssb_neigh",
    "ssb_error_string", "ssb_version",
]


class SsbArray(ctypes.Structure):
    _fields_ = [("ptrs", ctypes.c_void_p), ("base", ctypes.c_void_p), ("stride_bytes", ctypes.c_longlong)]


class SsbError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libseissol_b200.so; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SsbError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.ssb_error_string.restype = ctypes.c_char_p
    lib.ssb_version.restype = ctypes.c_char_p
    A, P, LL, I = SsbArray, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.ssb_create.argtypes = [ctypes.POINTER(P), I, I, I, I]
    lib.ssb_destroy.argtypes = [P]
    lib.ssb_set_layout.argtypes = [P, I, P, P]
    lib.ssb_layout_reference.argtypes = [P]
    lib.ssb_query.argtypes = [P, P]
    lib.ssb_ader.argtypes = [P, A, A, A, A, P, LL, P, P]
    lib.ssb_local.argtypes = [P, A, A, A, P, LL, P, P]
    lib.ssb_ader_local.argtypes = [P, A, A, A, A, P, P, LL, P, P]
    lib.ssb_taylor.argtypes = [P, A, A, P, LL, P]
    lib.ssb_neigh.argtypes = [P, A, P, P, A, A, P, P, LL, P, P]
    _lib = lib
    return lib


def _check(code):
    if code != 0:
        raise SsbError(f"seissol_b200 error {code}: {load_library().ssb_error_string(code).decode()}")


def strided(t):
    """torch tensor [n, ...] (contiguous, on the GPU) -> base + stride descriptor."""
    assert t.is_cuda and t.is_contiguous()
    stride = t.stride(0) * t.element_size() if t.dim() > 0 and t.shape[0] > 0 else 0
    return SsbArray(None, t.data_ptr(), stride)


def pointer_table(t):
    """torch int64 tensor of device pointers (the reference's `real**` tables)."""
    assert t.is_cuda and t.dtype.is_floating_point is False
    return SsbArray(t.data_ptr(), None, 0)


NONE = SsbArray(None, None, 0)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(stream):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


class Context:
    """One (order, precision, equation) configuration; mirrors the reference's
    compile-time configuration (src/Config.h.in)."""

    def __init__(self, order, real_size, equation=0, mechanisms=0, reference_layout=False):
        self.lib = load_library()
        self.handle = ctypes.c_void_p()
        _check(self.lib.ssb_create(ctypes.byref(self.handle), order, real_size, equation, mechanisms))
        self.order, self.real_size = order, real_size
        if reference_layout:
            _check(self.lib.ssb_layout_reference(self.handle))
        self._query()

    def _query(self):
        out = (ctypes.c_int * 24)()
        _check(self.lib.ssb_query(self.handle, out))
        self.B, self.ld, self.qsize, self.dersize = out[0], out[1], out[2], out[3]
        self.der_offset = [out[4 + d] for d in range(self.order)]
        self.der_ld = [out[12 + d] for d in range(self.order)]
        self.e_local, self.e_neigh = out[20], out[21]

    def __del__(self):
        try:
            if self.handle:
                self.lib.ssb_destroy(self.handle)
        except Exception:
            pass

    @staticmethod
    def _coeffs(c):
        a = np.ascontiguousarray(c, dtype=np.float64)
        return a, a.ctypes.data_as(ctypes.c_void_p)

    def ader(self, dofs, idofs, derivs, local_int, coeffs, n, perm=None, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_ader(self.handle, dofs, idofs, derivs if derivs is not None else NONE,
                                 local_int, c, n, _ptr(perm), _stream(stream)))

    def local(self, dofs, idofs, local_int, face_types, n, perm=None, stream=None):
        _check(self.lib.ssb_local(self.handle, dofs, idofs, local_int, _ptr(face_types), n, _ptr(perm),
                                  _stream(stream)))

    def ader_local(self, dofs, idofs, derivs, local_int, face_types, coeffs, n, perm=None, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_ader_local(self.handle, dofs, idofs, derivs if derivs is not None else NONE,
                                       local_int, _ptr(face_types), c, n, _ptr(perm), _stream(stream)))

    def taylor(self, derivs, idofs, coeffs, n, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_taylor(self.handle, derivs, idofs, c, n, _stream(stream)))

    def neigh(self, dofs, idofs, neigh_int, face_types, face_rel, n, nb_idx=None, nb_ptrs=None,
              perm=None, stream=None):
        _check(self.lib.ssb_neigh(self.handle, dofs, _ptr(nb_ptrs), _ptr(nb_idx), idofs, neigh_int,
                                  _ptr(face_types), _ptr(face_rel), n, _ptr(perm), _stream(stream)))
