"""ctypes binding of include/digiham_b200.h (harness glue; see package docstring)."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libdigiham_b200.so"

RRC_WIDE = 0
RRC_NARROW = 1
FMT_F32, FMT_S16 = 0, 1
SHARD_SCATTER = 1
OPT_DMR_LC_FEC = 1
PROTO_DMR, PROTO_YSF, PROTO_POCSAG, PROTO_NXDN, PROTO_DSTAR = 0, 1, 2, 3, 4


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
    L.dh_demod_create.argtypes = [c_void_pp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                  ctypes.c_int]
    L.dh_demod_reserve.argtypes = [ctypes.c_void_p, ctypes.c_size_t, c_void_pp, ctypes.POINTER(ctypes.c_size_t)]
    L.dh_demod_max_symbols.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.dh_demod_max_symbols.restype = ctypes.c_size_t
    L.dh_demod_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                   ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    L.dh_demod_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_demod_set_split.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dh_demod_kernels_per_call.argtypes = [ctypes.c_void_p]
    L.dh_pipe_demod.argtypes = [ctypes.c_void_p]
    L.dh_pipe_demod.restype = ctypes.c_void_p
    L.dh_demod_destroy.argtypes = [ctypes.c_void_p]
    L.dh_demod_destroy.restype = None
    L.dh_decoder_create.argtypes = [c_void_pp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int]
    L.dh_decoder_reserve.argtypes = [ctypes.c_void_p, ctypes.c_size_t, c_void_pp, ctypes.POINTER(ctypes.c_size_t)]
    L.dh_decoder_set_slot_filter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint8]
    L.dh_decoder_set_meta_kv.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dh_decoder_set_option.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.dh_decoder_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                     ctypes.c_size_t, ctypes.c_void_p]
    L.dh_decoder_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_decoder_output.argtypes = [ctypes.c_void_p, ctypes.c_uint32, c_void_pp, ctypes.POINTER(ctypes.c_size_t)]
    L.dh_decoder_meta.argtypes = [ctypes.c_void_p, ctypes.c_uint32, c_void_pp, ctypes.POINTER(ctypes.c_size_t)]
    L.dh_decoder_totals.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    L.dh_decoder_clear.argtypes = [ctypes.c_void_p]
    L.dh_decoder_destroy.argtypes = [ctypes.c_void_p]
    L.dh_decoder_destroy.restype = None
    L.dh_pipe_create.argtypes = [c_void_pp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_size_t]
    L.dh_pipe_process_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                         ctypes.c_void_p]
    L.dh_pipe_process_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                       ctypes.c_void_p]
    L.dh_pipe_host_pitch.argtypes = [ctypes.c_void_p]
    L.dh_pipe_host_pitch.restype = ctypes.c_size_t
    L.dh_pipe_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_pipe_decoder.argtypes = [ctypes.c_void_p]
    L.dh_pipe_decoder.restype = ctypes.c_void_p
    L.dh_pipe_last_symbols.argtypes = [ctypes.c_void_p, c_void_pp, ctypes.POINTER(ctypes.c_size_t), c_void_pp]
    L.dh_pipe_read_symbols.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.POINTER(ctypes.c_size_t)]
    L.dh_pipe_submit_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
    L.dh_pipe_collect_step.argtypes = [ctypes.c_void_p]
    L.dh_decoder_select_results.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dh_decoder_collect_results.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.dh_pipe_set_sub_chunk.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.dh_pipe_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dh_pipe_set_async.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.dh_pipe_sync.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_pipe_discard.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_pipe_stage_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    L.dh_pipe_launch_count.argtypes = [ctypes.c_void_p]
    L.dh_pipe_launch_count.restype = ctypes.c_uint64
    L.dh_decoder_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    L.dh_decoder_discard.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_pipe_destroy.argtypes = [ctypes.c_void_p]
    L.dh_pipe_destroy.restype = None
    for bank in ("rrc", "demod", "decoder", "pipe"):
        getattr(L, "dh_%s_state_size" % bank).argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        getattr(L, "dh_%s_state_export" % bank).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                            ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]
        getattr(L, "dh_%s_state_import" % bank).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                            ctypes.c_void_p]
    L.dh_test_fec.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    L.dh_test_bptc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    L.dh_test_viterbi.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
    L.dh_host_alloc.argtypes = [c_void_pp, ctypes.c_size_t, ctypes.c_int]
    L.dh_host_free.argtypes = [ctypes.c_void_p]
    L.dh_rrc_process_s16.argtypes = L.dh_rrc_process.argtypes
    L.dh_pipe_process_device_s16.argtypes = L.dh_pipe_process_device.argtypes
    L.dh_pipe_process_host_s16.argtypes = L.dh_pipe_process_host.argtypes
    L.dh_pipe_submit_host_s16.argtypes = L.dh_pipe_submit_host.argtypes
    L.dh_pipe_host_pitch_s16.argtypes = [ctypes.c_void_p]
    L.dh_pipe_host_pitch_s16.restype = ctypes.c_size_t
    L.dh_pipe_input_event.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    for name in ("dh_rrc_channels", "dh_demod_channels", "dh_decoder_channels", "dh_dvf_channels", "dh_pipe_channels"):
        getattr(L, name).argtypes = [ctypes.c_void_p]
        getattr(L, name).restype = ctypes.c_uint32
    c_u64_p = ctypes.POINTER(ctypes.c_uint64)
    L.dh_shard_unique_id.argtypes = [ctypes.c_void_p]
    L.dh_shard_comm_init.argtypes = [c_void_pp, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.dh_shard_comm_destroy.argtypes = [ctypes.c_void_p]
    L.dh_shard_channel_range.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, c_u64_p, c_u64_p]
    L.dh_shard_wire_layout.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32),
                                       ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_size_t)]
    L.dh_shard_create.argtypes = [c_void_pp, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_uint64, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
    L.dh_shard_pitch.argtypes = [ctypes.c_void_p]
    L.dh_shard_pitch.restype = ctypes.c_size_t
    L.dh_shard_local_channels.argtypes = [ctypes.c_void_p]
    L.dh_shard_local_channels.restype = ctypes.c_uint32
    L.dh_shard_pipe.argtypes = [ctypes.c_void_p]
    L.dh_shard_pipe.restype = ctypes.c_void_p
    L.dh_shard_submit_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_void_p]
    L.dh_shard_collect_step.argtypes = [ctypes.c_void_p]
    L.dh_shard_discard_step.argtypes = [ctypes.c_void_p]
    L.dh_shard_sync.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_shard_output.argtypes = [ctypes.c_void_p, ctypes.c_uint64, c_void_pp, ctypes.POINTER(ctypes.c_size_t)]
    L.dh_shard_meta.argtypes = [ctypes.c_void_p, ctypes.c_uint64, c_void_pp, ctypes.POINTER(ctypes.c_size_t)]
    L.dh_shard_clear.argtypes = [ctypes.c_void_p]
    L.dh_shard_scatter_path.argtypes = [ctypes.c_void_p]
    L.dh_shard_gather_path.argtypes = [ctypes.c_void_p]
    L.dh_shard_stats.argtypes = [ctypes.c_void_p, c_u64_p, c_u64_p, c_u64_p]
    L.dh_shard_destroy.argtypes = [ctypes.c_void_p]
    L.dh_shard_destroy.restype = None
    L.dh_dvf_create.argtypes = [c_void_pp, ctypes.c_int, ctypes.c_uint32]
    L.dh_dvf_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_void_p]
    L.dh_dvf_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dh_dvf_destroy.argtypes = [ctypes.c_void_p]
    L.dh_dvf_destroy.restype = None
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


class PinnedBlock:
    """A page-locked host block from dh_host_alloc, viewed as a torch tensor [rows, pitch] (`.tensor`)."""

    def __init__(self, rows, pitch, dtype=torch.float32, write_combined=False):
        self._p = ctypes.c_void_p()
        self.nbytes = int(rows) * int(pitch) * torch.empty((), dtype=dtype).element_size()
        check(lib().dh_host_alloc(ctypes.byref(self._p), self.nbytes, int(write_combined)))
        buf = (ctypes.c_char * self.nbytes).from_address(self._p.value)
        self.tensor = torch.frombuffer(buf, dtype=dtype).view(int(rows), int(pitch))

    def close(self):
        if self._p:
            self.tensor = None
            lib().dh_host_free(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _export_state(bank, handle, stream=None):
    n = ctypes.c_size_t()
    check(getattr(lib(), "dh_%s_state_size" % bank)(handle, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(n.value)
    w = ctypes.c_size_t()
    check(getattr(lib(), "dh_%s_state_export" % bank)(handle, buf, n.value, ctypes.byref(w), _stream_ptr(stream)))
    return buf.raw[:w.value]


def _import_state(bank, handle, blob, stream=None):
    check(getattr(lib(), "dh_%s_state_import" % bank)(handle, blob, len(blob), _stream_ptr(stream)))


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
        """x: float32 (or int16: fused `csdr convert -i s16 -o float`) CUDA tensor [channels, pitch]; filters the
        first n samples of every row."""
        assert x.is_cuda and x.dtype in (torch.float32, torch.int16) and x.dim() == 2 and x.shape[0] == self.channels
        assert x.stride(1) == 1
        if n is None:
            n = x.shape[1]
        if out is None:
            out = torch.empty((self.channels, pitch4(n)), dtype=torch.float32, device=x.device)
        fn = lib().dh_rrc_process_s16 if x.dtype == torch.int16 else lib().dh_rrc_process
        check(fn(self._h, x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), n, _stream_ptr(stream)))
        return out

    def reset(self, stream=None):
        check(lib().dh_rrc_reset(self._h, _stream_ptr(stream)))

    def export_state(self, stream=None):
        """The bank's per-channel state as an opaque bytes blob (dh_rrc_state_export)."""
        return _export_state("rrc", self._h, stream)

    def import_state(self, blob, stream=None):
        _import_state("rrc", self._h, blob, stream)

    def close(self):
        if self._h:
            lib().dh_rrc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DemodBank:
    """N x Digiham::Fsk::GfskDemodulator / FskDemodulator (reference include/gfsk_demodulator.hpp:12-33,
    include/fsk_demodulator.hpp:12-33) on one GPU."""

    def __init__(self, channels, sps=10, four_level=True, invert=False, device="cuda:0"):
        self._h = ctypes.c_void_p()
        self.channels = int(channels)
        self.sps = int(sps)
        self.device = torch.device(device)
        check(lib().dh_demod_create(ctypes.byref(self._h), _dev_index(device), self.channels, int(four_level),
                                    self.sps, int(invert)))

    def max_symbols(self, n):
        return lib().dh_demod_max_symbols(self._h, n)

    def set_split(self, enable):
        """dh_demod_set_split: one kernel (False), search chain + per-symbol + per-block kernels (True), or chosen
        per call from the bank and call size (None, the default)."""
        check(lib().dh_demod_set_split(self._h, -1 if enable is None else int(bool(enable))))

    @property
    def kernels_per_call(self):
        return lib().dh_demod_kernels_per_call(self._h)

    def reserve(self, max_n):
        """Returns (device pointer, pitch) of the zero-copy input block."""
        ptr = ctypes.c_void_p()
        pitch = ctypes.c_size_t()
        check(lib().dh_demod_reserve(self._h, max_n, ctypes.byref(ptr), ctypes.byref(pitch)))
        return ptr.value, pitch.value

    def process(self, x, n=None, sym=None, nsym=None, stream=None):
        """x: float32 CUDA tensor [channels, >=n] or a (ptr, pitch) pair from reserve()."""
        if isinstance(x, tuple):
            ptr, pitch = x
            assert n is not None
            dev = self.device
        else:
            assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] == self.channels
            ptr, pitch = x.data_ptr(), x.stride(0)
            dev = x.device
            if n is None:
                n = x.shape[1]
        if sym is None:
            sym = torch.empty((self.channels, self.max_symbols(n)), dtype=torch.uint8, device=dev)
        if nsym is None:
            nsym = torch.empty((self.channels,), dtype=torch.int32, device=dev)
        check(lib().dh_demod_process(self._h, ptr, pitch, n, sym.data_ptr(), sym.stride(0), nsym.data_ptr(),
                                     _stream_ptr(stream)))
        return sym, nsym

    def reset(self, stream=None):
        check(lib().dh_demod_reset(self._h, _stream_ptr(stream)))

    def export_state(self, stream=None):
        """The bank's per-channel state as an opaque bytes blob (dh_demod_state_export)."""
        return _export_state("demod", self._h, stream)

    def import_state(self, blob, stream=None):
        _import_state("demod", self._h, blob, stream)

    def close(self):
        if self._h:
            lib().dh_demod_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DecoderBank:
    """N x Digiham::{Dmr,Ysf,Pocsag}::Decoder (reference include/decoder.hpp:17-30) on one GPU."""

    def __init__(self, channels, proto=PROTO_DMR, device="cuda:0"):
        self._h = ctypes.c_void_p()
        self.channels = int(channels)
        self.proto = proto
        self.device = torch.device(device)
        check(lib().dh_decoder_create(ctypes.byref(self._h), _dev_index(device), self.channels, proto))

    def reserve(self, max_syms):
        ptr = ctypes.c_void_p()
        pitch = ctypes.c_size_t()
        check(lib().dh_decoder_reserve(self._h, max_syms, ctypes.byref(ptr), ctypes.byref(pitch)))
        return ptr.value, pitch.value

    def set_slot_filter(self, filt, channel=-1):
        check(lib().dh_decoder_set_slot_filter(self._h, channel, filt))

    def set_option(self, option, value, channel=-1):
        """Opt-in modes beyond the reference (dh_decoder_set_option), e.g. OPT_DMR_LC_FEC."""
        check(lib().dh_decoder_set_option(self._h, channel, option, value))

    def process(self, sym, nsym, max_nsym=None, stream=None):
        """sym: uint8 CUDA tensor [channels, >=max_nsym] or (ptr, pitch); nsym: int32 CUDA tensor [channels]."""
        if isinstance(sym, tuple):
            ptr, pitch = sym
            assert max_nsym is not None
        else:
            assert sym.is_cuda and sym.dtype == torch.uint8 and sym.dim() == 2 and sym.shape[0] == self.channels
            ptr, pitch = sym.data_ptr(), sym.stride(0)
            if max_nsym is None:
                max_nsym = sym.shape[1]
        assert nsym.is_cuda and nsym.dtype == torch.int32 and nsym.numel() == self.channels
        check(lib().dh_decoder_process(self._h, ptr, pitch, nsym.data_ptr(), max_nsym, _stream_ptr(stream)))

    def collect(self, stream=None):
        check(lib().dh_decoder_collect(self._h, _stream_ptr(stream)))

    def output(self, channel):
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        check(lib().dh_decoder_output(self._h, channel, ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value) if n.value else b""

    def meta(self, channel):
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        check(lib().dh_decoder_meta(self._h, channel, ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value) if n.value else b""

    def totals(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().dh_decoder_totals(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def stats(self):
        """(events replayed, bytes copied device->host) since creation."""
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().dh_decoder_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def discard(self, stream=None):
        check(lib().dh_decoder_discard(self._h, _stream_ptr(stream)))

    def clear(self):
        check(lib().dh_decoder_clear(self._h))

    def export_state(self, stream=None):
        """The bank's per-channel state as an opaque bytes blob (dh_decoder_state_export)."""
        return _export_state("decoder", self._h, stream)

    def import_state(self, blob, stream=None):
        _import_state("decoder", self._h, blob, stream)

    def close(self):
        if self._h:
            lib().dh_decoder_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipe:
    """rrc_filter | gfsk_demodulator | <proto>_decoder for N channels (reference examples/*-decoder.sh)."""

    def __init__(self, channels, proto=PROTO_DMR, max_chunk=48000, device="cuda:0"):
        self._h = ctypes.c_void_p()
        self.channels = int(channels)
        self.proto = proto
        self.max_chunk = int(max_chunk)
        self.device = torch.device(device)
        check(lib().dh_pipe_create(ctypes.byref(self._h), _dev_index(device), self.channels, proto, self.max_chunk))
        # a non-owning view of the pipe's decoder bank
        self.decoder = DecoderBank.__new__(DecoderBank)
        self.decoder._h = ctypes.c_void_p(lib().dh_pipe_decoder(self._h))
        self.decoder.channels = self.channels
        self.decoder.proto = proto
        self.decoder.device = self.device
        self.decoder.close = lambda: None

    def set_demod_split(self, enable):
        """Schedule of the pipe's demodulator bank (dh_demod_set_split on dh_pipe_demod)."""
        check(lib().dh_demod_set_split(ctypes.c_void_p(lib().dh_pipe_demod(self._h)),
                                       -1 if enable is None else int(bool(enable))))

    @property
    def demod_kernels_per_call(self):
        return lib().dh_demod_kernels_per_call(ctypes.c_void_p(lib().dh_pipe_demod(self._h)))

    @property
    def host_pitch(self):
        return lib().dh_pipe_host_pitch(self._h)

    @property
    def host_pitch_s16(self):
        return lib().dh_pipe_host_pitch_s16(self._h)

    def process(self, x, n=None, stream=None):
        """x: float32 or int16 tensor [channels, pitch]; CUDA tensors are consumed in place, CPU tensors are copied."""
        assert x.dtype in (torch.float32, torch.int16) and x.dim() == 2 and x.shape[0] == self.channels
        assert x.stride(1) == 1
        if n is None:
            n = x.shape[1]
        s16 = x.dtype == torch.int16
        L = lib()
        if x.is_cuda:
            fn = L.dh_pipe_process_device_s16 if s16 else L.dh_pipe_process_device
        else:
            fn = L.dh_pipe_process_host_s16 if s16 else L.dh_pipe_process_host
        check(fn(self._h, x.data_ptr(), x.stride(0), n, _stream_ptr(stream)))

    def collect(self, stream=None):
        check(lib().dh_pipe_collect(self._h, _stream_ptr(stream)))

    def submit(self, x, n=None):
        """Streaming interface: asynchronous upload + kernels of one block from PINNED host memory."""
        assert (not x.is_cuda) and x.is_pinned() and x.dtype in (torch.float32, torch.int16)
        assert x.shape[0] == self.channels
        if n is None:
            n = x.shape[1]
        fn = lib().dh_pipe_submit_host_s16 if x.dtype == torch.int16 else lib().dh_pipe_submit_host
        check(fn(self._h, x.data_ptr(), x.stride(0), n))

    def collect_step(self):
        """Waits for the oldest submitted step and appends its results to the per-channel host buffers."""
        check(lib().dh_pipe_collect_step(self._h))

    def set_sub_chunk(self, sub_chunk):
        """Software pipelining granularity inside one process call (0 = three kernels back to back)."""
        check(lib().dh_pipe_set_sub_chunk(self._h, int(sub_chunk)))

    def set_profiling(self, enable):
        check(lib().dh_pipe_set_profiling(self._h, int(enable)))

    def set_async(self, enable, stream=None):
        """Cross-call software pipelining: K1 of call i+1 overlaps K2 + decoder of call i (join with sync())."""
        check(lib().dh_pipe_set_async(self._h, int(enable), _stream_ptr(stream)))

    def sync(self, stream=None):
        check(lib().dh_pipe_sync(self._h, _stream_ptr(stream)))

    def discard(self, stream=None):
        check(lib().dh_pipe_discard(self._h, _stream_ptr(stream)))

    def stage_times(self):
        """([ms_rrc, ms_demod, ms_decoder] summed, calls) since the previous query (synchronises)."""
        ms = (ctypes.c_double * 3)()
        calls = ctypes.c_uint64()
        check(lib().dh_pipe_stage_times(self._h, ms, ctypes.byref(calls)))
        return [ms[0], ms[1], ms[2]], calls.value

    @property
    def launch_count(self):
        return lib().dh_pipe_launch_count(self._h)

    def last_symbols(self, channel):
        """numpy uint8 array: the symbols the demodulator emitted for `channel` in the last process call."""
        import numpy as np
        buf = np.empty(self.max_chunk // 4 + 4096, dtype=np.uint8)
        n = ctypes.c_size_t()
        check(lib().dh_pipe_read_symbols(self._h, channel, buf.ctypes.data, buf.size, ctypes.byref(n)))
        return buf[:n.value].copy()

    def output(self, channel):
        return self.decoder.output(channel)

    def meta(self, channel):
        return self.decoder.meta(channel)

    def totals(self):
        return self.decoder.totals()

    def export_state(self, stream=None):
        """The bank's per-channel state as an opaque bytes blob (dh_pipe_state_export)."""
        return _export_state("pipe", self._h, stream)

    def import_state(self, blob, stream=None):
        _import_state("pipe", self._h, blob, stream)

    def close(self):
        if self._h:
            lib().dh_pipe_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DvfBank:
    """N x Digiham::DigitalVoice::DigitalVoiceFilter (reference include/digitalvoice_filter.hpp:12-19)."""

    def __init__(self, channels, device="cuda:0"):
        self._h = ctypes.c_void_p()
        self.channels = int(channels)
        check(lib().dh_dvf_create(ctypes.byref(self._h), _dev_index(device), self.channels))

    def process(self, x, out=None, n=None, stream=None):
        assert x.is_cuda and x.dtype == torch.int16 and x.dim() == 2 and x.shape[0] == self.channels
        if n is None:
            n = x.shape[1]
        if out is None:
            out = torch.empty_like(x)
        check(lib().dh_dvf_process(self._h, x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), n,
                                   _stream_ptr(stream)))
        return out

    def reset(self, stream=None):
        check(lib().dh_dvf_reset(self._h, _stream_ptr(stream)))

    def close(self):
        if self._h:
            lib().dh_dvf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
