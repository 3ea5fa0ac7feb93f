"""ctypes binding of include/digiham_b200.h (harness glue; see package docstring)."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libdigiham_b200.so"

RRC_WIDE = 0
RRC_NARROW = 1


class DhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("digiham_b200 error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    return os.path.join(_HERE, _LIB_NAME)


_lib = None


def lib():
    """Load libdigiham_b200.so.  Fails loudly if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C digiham_b200/csrc` (the CUDA path is the only path)" % path
        )
    L = ctypes.CDLL(path)
    c_void_pp = ctypes.POINTER(ctypes.c_void_p)
    L.dh_last_error.restype = ctypes.c_char_p
    L.dh_version.restype = ctypes.c_char_p
    L.dh_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.dh_rrc_create.argtypes = [c_void_pp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int]
    L.dh_rrc_create_custom.argtypes = [c_void_pp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double,
                                       ctypes.c_void_p]
    L.dh_rrc_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_void_p]
    L.dh_rrc_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_rrc_destroy.argtypes = [ctypes.c_void_p]
    L.dh_rrc_destroy.restype = None
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise DhError(rc, lib().dh_last_error().decode("utf-8", "replace"))


def _stream_ptr(stream):
    if stream is None:
        stream = torch.cuda.current_stream()
    return ctypes.c_void_p(stream.cuda_stream)


def _dev_index(device):
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def pitch4(n):
    return (int(n) + 3) & ~3


class RrcBank:
    """N x Digiham::RrcFilter::RrcFilter (reference include/rrc_filter.hpp:10-31) on one GPU."""

    def __init__(self, channels, kind=RRC_WIDE, device="cuda:0", custom=None):
        self._h = ctypes.c_void_p()
        self.channels = int(channels)
        self.device = torch.device(device)
        if custom is None:
            check(lib().dh_rrc_create(ctypes.byref(self._h), _dev_index(device), self.channels, kind))
        else:
            import numpy as np
            n_zeros, gain, coeffs = custom
            c = np.ascontiguousarray(coeffs, dtype=np.float32)
            assert c.size == n_zeros + 1
            check(lib().dh_rrc_create_custom(ctypes.byref(self._h), _dev_index(device), self.channels, n_zeros,
                                             float(gain), c.ctypes.data))

    def process(self, x, out=None, n=None, stream=None):
        """x: float32 CUDA tensor [channels, pitch]; filters the first n samples of every row."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] == self.channels
        assert x.stride(1) == 1
        if n is None:
            n = x.shape[1]
        if out is None:
            out = torch.empty((self.channels, pitch4(n)), dtype=torch.float32, device=x.device)
        check(lib().dh_rrc_process(self._h, x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), n,
                                   _stream_ptr(stream)))
        return out

    def reset(self, stream=None):
        check(lib().dh_rrc_reset(self._h, _stream_ptr(stream)))

    def close(self):
        if self._h:
            lib().dh_rrc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
