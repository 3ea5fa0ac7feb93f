"""D-Star parity: dh_decoder_* (DH_PROTO_DSTAR) and the fsk -s 10 | dstar pipe vs the CPU
oracle: voice byte stream byte-exact, metadata lines string-exact.  Streams contain radio headers (voice, data,
uncorrectable), late entry on the voice sync, slow data (20-character message, header resend, DPRS and NMEA GGA
sentences), both terminator forms, noise gaps and bit errors."""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

ERRS = [0.0, 0.001, 0.005, 0.02, 0.05]


def _streams(C, frames, seed):
    out = [synth.dstar_symbols(frames, seed=seed * 100 + ch, bit_errors=ERRS[ch % len(ERRS)]) for ch in range(C)]
    n = min(len(s) for s in out)
    return np.stack([s[:n] for s in out])


def _check(bank, sym, chunk=0):
    orc = oracle_lib.best()
    total = 0
    for ch in range(sym.shape[0]):
        ref_out, ref_meta = orc.decode(oracle_lib.PROTO_DSTAR, sym[ch], chunk=chunk)
        assert bank.output(ch) == ref_out.tobytes(), "channel %d bytes differ (%d vs %d)" % (
            ch, len(bank.output(ch)), ref_out.size)
        assert bank.meta(ch) == ref_meta, "channel %d meta differs:\n%s\n---\n%s" % (
            ch, bank.meta(ch).decode(errors="replace")[:800], ref_meta.decode(errors="replace")[:800])
        total += ref_out.size + len(ref_meta)
    return total


def test_dstar_decoder_whole_stream():
    import digiham_b200 as dh
    C = 40
    sym = _streams(C, 300, seed=1)
    bank = dh.DecoderBank(C, dh.PROTO_DSTAR)
    bank.process(torch.from_numpy(sym).cuda(), torch.full((C,), sym.shape[1], dtype=torch.int32, device="cuda"))
    bank.collect()
    assert _check(bank, sym) > 40000
    bank.close()


def test_dstar_decoder_noise_and_four_level_bytes():
    """random bits (spurious syncs, rejected headers) and random symbols 0..3 (bit 1 counts in the distances)"""
    import digiham_b200 as dh
    C = 64
    rng = np.random.default_rng(5)
    sym = rng.integers(0, 2, size=(C, 120000)).astype(np.uint8)
    sym[C // 2:] = rng.integers(0, 4, size=(C - C // 2, 120000)).astype(np.uint8)
    bank = dh.DecoderBank(C, dh.PROTO_DSTAR)
    bank.process(torch.from_numpy(sym).cuda(), torch.full((C,), sym.shape[1], dtype=torch.int32, device="cuda"))
    bank.collect()
    _check(bank, sym)
    bank.close()


def test_dstar_decoder_streaming_chunks():
    import digiham_b200 as dh
    C = 10
    sym = _streams(C, 200, seed=2)
    n = sym.shape[1]
    bank = dh.DecoderBank(C, dh.PROTO_DSTAR)
    rng = np.random.default_rng(3)
    pos = 0
    while pos < n:
        c = int(min(n - pos, rng.choice([1, 23, 24, 25, 95, 96, 97, 119, 120, 121, 659, 660, 661, 4800])))
        bank.process(torch.from_numpy(np.ascontiguousarray(sym[:, pos:pos + c])).cuda(),
                     torch.full((C,), c, dtype=torch.int32, device="cuda"))
        bank.collect()
        pos += c
    _check(bank, sym, chunk=128)
    bank.close()


def test_dstar_pipe_vs_oracle():
    """fsk_demodulator -s 10 | dstar_decoder (examples/dstar-decoder.sh:19-21)."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 16
    sym = _streams(C, 120, seed=4)
    # a preamble of alternating bits lets the demodulator's level tracking settle before the first header
    sym = np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), (C, 150)), sym], axis=1)
    n = sym.shape[1] * 10
    x = synth.modulate_batch(sym, n, sps=10, levels=synth.LEVELS2, amplitude=0.5,
                             ppm=np.array([0, 30, -30, 60] * 4, dtype=np.float64),
                             phase=np.arange(C, dtype=np.float64) * 3, snr_db=np.array([np.inf, 20, 14, 9] * 4),
                             seed=9, device="cuda")
    pipe = dh.Pipe(C, dh.PROTO_DSTAR, max_chunk=48000)
    for pos in range(0, n, 48000):
        c = min(48000, n - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = x[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()
    xc = x[:, :n].cpu().numpy()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DSTAR, xc, threads=8, meta_cap=1 << 15)
    total = 0
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes(), ch
        assert pipe.meta(ch) == metas[ch], ch
        total += outs[ch].size + len(metas[ch])
    assert total > 5000
    pipe.close()
