"""digiham_b200 — B200-native hot path of jketterl/digiham behind a C ABI.

The product is ``libdigiham_b200.so`` (hand-written sm_100a CUDA + an ``extern "C"`` layer declared in
``include/digiham_b200.h``) and the header-compatible C++ facade classes in ``include/*.hpp``.  This Python
package is harness glue only: a ctypes binding of the C ABI used by ``tests/``, ``bench.py`` and
``__graft_entry__.py``; PyTorch merely owns device buffers, streams and the process group.

There is no CPU fallback: importing works anywhere, but every bank constructor raises without a CUDA device
or without the compiled library.
"""
from ._capi import (  # noqa: F401
    DhError,
    lib,
    lib_path,
    RrcBank,
    DemodBank,
    DecoderBank,
    Pipe,
    DvfBank,
    PinnedBlock,
    PROTO_DMR,
    PROTO_YSF,
    PROTO_POCSAG,
    PROTO_NXDN,
    PROTO_DSTAR,
    RRC_WIDE,
    RRC_NARROW,
    FMT_F32,
    FMT_S16,
    SHARD_SCATTER,
    OPT_DMR_LC_FEC,
)

__all__ = ["DhError", "lib", "lib_path", "RrcBank", "DemodBank", "DecoderBank", "Pipe", "DvfBank", "PinnedBlock", "PROTO_DMR", "PROTO_YSF", "PROTO_POCSAG", "PROTO_NXDN", "PROTO_DSTAR", "RRC_WIDE", "RRC_NARROW", "FMT_F32", "FMT_S16", "SHARD_SCATTER", "OPT_DMR_LC_FEC"]
