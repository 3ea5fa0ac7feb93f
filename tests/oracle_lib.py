"""ctypes access to the CPU checkers behind oracle/oracle_api.h (TEST INFRASTRUCTURE).

``ref()``  -> oracle/_ref/libdigiham_ref.so  (unmodified reference sources; built by oracle/Makefile where
              /root/reference exists, shipped prebuilt to the GPU box)
``port()`` -> oracle/liboracle_port.so       (independent restatement, always buildable)
``best()`` -> the reference build when present, else the port.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libdigiham_ref.so")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle_port.so")

PROTO_DMR, PROTO_YSF, PROTO_POCSAG, PROTO_NXDN, PROTO_DSTAR = 0, 1, 2, 3, 4
FEC_NAMES = ["hamming_7_4", "hamming_13_9", "hamming_15_11", "hamming_16_11", "qr_16_7", "golay_20_8",
             "golay_24_12", "bch_31_21"]
FEC_BITS = [7, 13, 15, 16, 16, 20, 24, 31]
FEC_PARITY_BITS = [3, 4, 4, 5, 9, 12, 12, 10]

_sz = ctypes.c_size_t
_vp = ctypes.c_void_p


def _bind(path):
    L = ctypes.CDLL(path)
    L.orc_kind.restype = ctypes.c_char_p
    L.orc_rrc.restype = _sz
    L.orc_rrc.argtypes = [ctypes.c_int, _vp, _sz, _sz, _vp]
    L.orc_demod.restype = _sz
    L.orc_demod.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.c_int, _vp, _sz, _sz, _vp, _sz]
    L.orc_decode.restype = _sz
    L.orc_decode.argtypes = [ctypes.c_int, _vp, _sz, _sz, ctypes.c_int, _vp, _sz, _vp, _sz, ctypes.POINTER(_sz)]
    L.orc_pipe.restype = _sz
    L.orc_pipe.argtypes = [ctypes.c_int, _vp, _sz, _sz, ctypes.c_int, _vp, _sz, ctypes.POINTER(_sz), _vp, _sz, _vp,
                           _sz, ctypes.POINTER(_sz)]
    L.orc_pipe_batch.restype = _sz
    L.orc_pipe_batch.argtypes = [ctypes.c_int, _vp, _sz, _sz, _sz, ctypes.c_int, ctypes.c_int, _vp, _sz, _vp, _vp,
                                 _sz, _vp, _vp, _sz, _vp]
    L.orc_dvf.restype = _sz
    L.orc_dvf.argtypes = [_vp, _sz, _sz, _vp]
    L.orc_fec.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    L.orc_fec_syndrome.restype = ctypes.c_uint32
    L.orc_fec_syndrome.argtypes = [ctypes.c_int, ctypes.c_uint32]
    L.orc_bptc_196_96.argtypes = [_vp, _vp]
    L.orc_trellis.restype = ctypes.c_uint
    L.orc_trellis.argtypes = [_vp, ctypes.c_uint, _vp]
    L.orc_crc16.restype = ctypes.c_uint16
    L.orc_crc16.argtypes = [_vp, ctypes.c_int]
    L.orc_whitening.argtypes = [_vp, _vp, ctypes.c_uint]
    L.orc_whitening.restype = None
    L.orc_hamming_distance.restype = ctypes.c_uint
    L.orc_hamming_distance.argtypes = [_vp, _vp, _sz]
    if hasattr(L, "orc_fec_batch"):
        L.orc_fec_batch.restype = None
        L.orc_fec_batch.argtypes = [ctypes.c_int, _vp, _vp, _sz, ctypes.c_int]
        L.orc_bptc_batch.restype = None
        L.orc_bptc_batch.argtypes = [_vp, _vp, _vp, _sz, ctypes.c_int]
        L.orc_trellis_batch.restype = None
        L.orc_trellis_batch.argtypes = [ctypes.c_int, _vp, _sz, ctypes.c_uint, _vp, _sz, _vp, _sz, ctypes.c_int]
    if hasattr(L, "orc_nxdn_trellis"):
        L.orc_nxdn_trellis.restype = ctypes.c_uint
        L.orc_nxdn_trellis.argtypes = [_vp, ctypes.c_uint, _vp]
        L.orc_nxdn_sacch.argtypes = [_vp, _vp]
        L.orc_nxdn_facch1.argtypes = [_vp]
    if hasattr(L, "orc_dstar_header"):
        L.orc_dstar_header.argtypes = [_vp, _vp, _sz]
    return L


class Oracle:
    def __init__(self, path):
        self.path = path
        self.L = _bind(path)
        self.kind = self.L.orc_kind().decode()

    def rrc(self, x, narrow=False, chunk=0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        n = self.L.orc_rrc(int(narrow), x.ctypes.data, x.size, chunk, out.ctypes.data)
        assert n == x.size
        return out

    def demod(self, x, sps=10, four_level=True, invert=False, chunk=0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        cap = x.size // max(1, sps - 1) + 16
        out = np.empty(cap, dtype=np.uint8)
        n = self.L.orc_demod(int(four_level), sps, int(invert), x.ctypes.data, x.size, chunk, out.ctypes.data, cap)
        assert n <= cap
        return out[:n].copy()

    def decode(self, proto, sym, chunk=0, slot_filter=3):
        sym = np.ascontiguousarray(sym, dtype=np.uint8)
        out_cap = sym.size + 4096
        meta_cap = 1 << 20
        out = np.empty(out_cap, dtype=np.uint8)
        meta = ctypes.create_string_buffer(meta_cap)
        ml = _sz(0)
        n = self.L.orc_decode(proto, sym.ctypes.data, sym.size, chunk, slot_filter, out.ctypes.data, out_cap, meta,
                              meta_cap, ctypes.byref(ml))
        assert n <= out_cap
        return out[:n].copy(), meta.raw[:ml.value]

    def pipe(self, proto, x, chunk=0, slot_filter=3):
        x = np.ascontiguousarray(x, dtype=np.float32)
        sym_cap = x.size // 8 + 64
        out_cap = x.size // 8 + 4096
        meta_cap = 1 << 20
        sym = np.empty(sym_cap, dtype=np.uint8)
        out = np.empty(out_cap, dtype=np.uint8)
        meta = ctypes.create_string_buffer(meta_cap)
        ns, ml = _sz(0), _sz(0)
        n = self.L.orc_pipe(proto, x.ctypes.data, x.size, chunk, slot_filter, sym.ctypes.data, sym_cap,
                            ctypes.byref(ns), out.ctypes.data, out_cap, meta, meta_cap, ctypes.byref(ml))
        return sym[:ns.value].copy(), out[:n].copy(), meta.raw[:ml.value]

    def pipe_batch(self, proto, x, threads=1, chunk=0, slot_filter=3, want_sym=False, meta_cap=4096):
        """x: [nch, n] float32.  Returns (sym list|None, out list, meta list)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        nch, n = x.shape
        sym_cap = n // 8 + 64
        out_cap = n // 8 + 4096
        sym = np.empty((nch, sym_cap), dtype=np.uint8) if want_sym else None
        nsym = np.zeros(nch, dtype=np.uint64)
        out = np.empty((nch, out_cap), dtype=np.uint8)
        olen = np.zeros(nch, dtype=np.uint64)
        meta = np.zeros((nch, meta_cap), dtype=np.uint8)
        mlen = np.zeros(nch, dtype=np.uint64)
        self.L.orc_pipe_batch(proto, x.ctypes.data, nch, n, chunk, slot_filter, threads,
                              sym.ctypes.data if want_sym else None, sym_cap, nsym.ctypes.data,
                              out.ctypes.data, out_cap, olen.ctypes.data, meta.ctypes.data, meta_cap, mlen.ctypes.data)
        syms = [sym[c, :int(nsym[c])].copy() for c in range(nch)] if want_sym else None
        outs = [out[c, :int(olen[c])].copy() for c in range(nch)]
        metas = [meta[c, :int(mlen[c])].tobytes() for c in range(nch)]
        return syms, outs, metas

    def dvf(self, x, chunk=0):
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.empty_like(x)
        n = self.L.orc_dvf(x.ctypes.data, x.size, chunk, out.ctypes.data)
        assert n == x.size
        return out

    def fec(self, code, word):
        w = ctypes.c_uint32(int(word))
        ok = self.L.orc_fec(code, ctypes.byref(w))
        return bool(ok), w.value

    def fec_batch(self, code, words, threads=8):
        """words: uint32 array -> (ok uint8 array, corrected uint32 array)."""
        w = np.ascontiguousarray(words, dtype=np.uint32).copy()
        ok = np.zeros(w.size, dtype=np.uint8)
        self.L.orc_fec_batch(code, w.ctypes.data, ok.ctypes.data, w.size, threads)
        return ok, w

    def bptc_batch(self, payloads, threads=8):
        """payloads: [n, 25] uint8 -> (ok [n], out [n, 12])."""
        p = np.ascontiguousarray(payloads, dtype=np.uint8)
        out = np.zeros((p.shape[0], 12), dtype=np.uint8)
        ok = np.zeros(p.shape[0], dtype=np.uint8)
        self.L.orc_bptc_batch(p.ctypes.data, out.ctypes.data, ok.ctypes.data, p.shape[0], threads)
        return ok, out

    def trellis_batch(self, dibits, nxdn=False, threads=8):
        """dibits: [n, steps] one dibit per byte -> (metric [n], decoded bytes [n, ceil(steps / 8)], MSB first)."""
        d = np.ascontiguousarray(dibits, dtype=np.uint8) & 3
        n, steps = d.shape
        pad = (-steps) % 4
        dp = np.concatenate([d, np.zeros((n, pad), dtype=np.uint8)], axis=1).reshape(n, -1, 4)
        packed = np.ascontiguousarray((dp[:, :, 0] << 6) | (dp[:, :, 1] << 4) | (dp[:, :, 2] << 2) | dp[:, :, 3]).astype(np.uint8)
        out_stride = (steps + 7) // 8 + 2
        out = np.zeros((n, out_stride), dtype=np.uint8)
        metric = np.zeros(n, dtype=np.uint32)
        self.L.orc_trellis_batch(int(nxdn), packed.ctypes.data, packed.shape[1], steps, out.ctypes.data, out_stride,
                                 metric.ctypes.data, n, threads)
        return metric, out[:, :(steps + 7) // 8]

    def fec_syndrome(self, code, word):
        return self.L.orc_fec_syndrome(code, int(word))

    def bptc(self, payload25):
        p = np.ascontiguousarray(payload25, dtype=np.uint8)
        out = np.zeros(12, dtype=np.uint8)
        ok = self.L.orc_bptc_196_96(p.ctypes.data, out.ctypes.data)
        return bool(ok), out

    def trellis(self, packed, steps):
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.zeros((steps + 7) // 8, dtype=np.uint8)
        metric = self.L.orc_trellis(p.ctypes.data, steps, out.ctypes.data)
        return metric, out

    def nxdn_trellis(self, packed, nbits):
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.zeros((nbits + 15) // 16, dtype=np.uint8)
        metric = self.L.orc_nxdn_trellis(p.ctypes.data, nbits, out.ctypes.data)
        return metric, out

    def nxdn_sacch(self, dibits30):
        d = np.ascontiguousarray(dibits30, dtype=np.uint8)
        out = np.zeros(5, dtype=np.uint8)
        ok = self.L.orc_nxdn_sacch(d.ctypes.data, out.ctypes.data)
        return bool(ok), out

    def nxdn_facch1(self, dibits72):
        d = np.ascontiguousarray(dibits72, dtype=np.uint8)
        return self.L.orc_nxdn_facch1(d.ctypes.data)

    def dstar_header(self, bits660):
        d = np.ascontiguousarray(bits660, dtype=np.uint8)
        text = ctypes.create_string_buffer(512)
        rc = self.L.orc_dstar_header(d.ctypes.data, text, 512)
        return rc, text.value

    def crc16(self, data):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        return self.L.orc_crc16(d.ctypes.data, d.size)

    def whitening(self, data, nbits):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros((nbits + 7) // 8, dtype=np.uint8)
        self.L.orc_whitening(d.ctypes.data, out.ctypes.data, nbits)
        return out


_cache = {}


def _make(target):
    subprocess.run(["make", "-C", ORACLE_DIR, target], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.PIPE)


def have_reference_tree():
    return os.path.isdir(os.environ.get("DIGIHAM_REF_DIR", "/root/reference"))


def ref():
    """The compiled reference, or None when neither a prebuilt .so nor the reference tree exists."""
    if "ref" not in _cache:
        if not os.path.exists(REF_SO) and have_reference_tree():
            _make("ref")
        _cache["ref"] = Oracle(REF_SO) if os.path.exists(REF_SO) else None
    return _cache["ref"]


def port():
    if "port" not in _cache:
        _make("port")
        _cache["port"] = Oracle(PORT_SO)
    return _cache["port"]


def best():
    return ref() or port()
