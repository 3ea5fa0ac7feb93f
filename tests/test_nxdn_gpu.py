"""NXDN parity: dh_decoder_* (DH_PROTO_NXDN) and the rrc -n | gfsk -s 20 | nxdn pipe vs the
CPU oracle: voice byte stream byte-exact, metadata lines string-exact.  Streams contain calls with SACCH
superframes (VCALL: call type / source / destination), FACCH1-stolen halves (IDLE, other types, TX_RELEASE),
frames with broken LICH parity, non-superframe SACCH, UDCH and RCCH frames, noise gaps and symbol errors, so the
punctured Viterbi + CRC6/CRC12 paths succeed, correct and fail like the reference's."""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

ERRS = [0.0, 0.002, 0.01, 0.03, 0.08]


def _streams(C, frames, seed):
    out = [synth.nxdn_symbols(frames, seed=seed * 100 + ch, symbol_errors=ERRS[ch % len(ERRS)]) for ch in range(C)]
    n = min(len(s) for s in out)
    return np.stack([s[:n] for s in out])


def _check(bank, sym, chunk=0):
    orc = oracle_lib.best()
    total = 0
    for ch in range(sym.shape[0]):
        ref_out, ref_meta = orc.decode(oracle_lib.PROTO_NXDN, sym[ch], chunk=chunk)
        assert bank.output(ch) == ref_out.tobytes(), "channel %d bytes differ (%d vs %d)" % (
            ch, len(bank.output(ch)), ref_out.size)
        assert bank.meta(ch) == ref_meta, "channel %d meta differs:\n%s\n---\n%s" % (
            ch, bank.meta(ch).decode(errors="replace")[:500], ref_meta.decode(errors="replace")[:500])
        total += ref_out.size + len(ref_meta)
    return total


def test_nxdn_decoder_whole_stream():
    import digiham_b200 as dh
    C = 40
    sym = _streams(C, 150, seed=1)
    bank = dh.DecoderBank(C, dh.PROTO_NXDN)
    bank.process(torch.from_numpy(sym).cuda(), torch.full((C,), sym.shape[1], dtype=torch.int32, device="cuda"))
    bank.collect()
    assert _check(bank, sym) > 20000
    bank.close()


def test_nxdn_decoder_random_symbols():
    """pure noise: sync search, spurious syncs, LICH parity passes by chance, CRC failures"""
    import digiham_b200 as dh
    C = 64
    sym = np.random.default_rng(5).integers(0, 4, size=(C, 60000)).astype(np.uint8)
    bank = dh.DecoderBank(C, dh.PROTO_NXDN)
    bank.process(torch.from_numpy(sym).cuda(), torch.full((C,), sym.shape[1], dtype=torch.int32, device="cuda"))
    bank.collect()
    _check(bank, sym)
    bank.close()


def test_nxdn_decoder_streaming_chunks():
    import digiham_b200 as dh
    C = 10
    sym = _streams(C, 80, seed=2)
    n = sym.shape[1]
    bank = dh.DecoderBank(C, dh.PROTO_NXDN)
    rng = np.random.default_rng(3)
    pos = 0
    while pos < n:
        c = int(min(n - pos, rng.choice([1, 9, 10, 11, 47, 191, 192, 193, 1000, 2400])))
        bank.process(torch.from_numpy(np.ascontiguousarray(sym[:, pos:pos + c])).cuda(),
                     torch.full((C,), c, dtype=torch.int32, device="cuda"))
        bank.collect()
        pos += c
    _check(bank, sym, chunk=128)
    bank.close()


def test_nxdn_pipe_vs_oracle():
    """rrc_filter -n | gfsk_demodulator -s 20 | nxdn_decoder (examples/nxdn48-decoder.sh:19-23)."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 16
    sym = _streams(C, 40, seed=4)
    n = sym.shape[1] * 20
    x = synth.modulate_batch(sym, n, sps=20, amplitude=0.5, ppm=np.array([0, 30, -30, 60] * 4, dtype=np.float64),
                             phase=np.arange(C, dtype=np.float64) * 5, snr_db=np.array([np.inf, 20, 14, 9] * 4),
                             seed=9, device="cuda")
    pipe = dh.Pipe(C, dh.PROTO_NXDN, max_chunk=48000)
    for pos in range(0, n, 48000):
        c = min(48000, n - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = x[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()
    xc = x[:, :n].cpu().numpy()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_NXDN, xc, threads=8, meta_cap=1 << 15)
    total = 0
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes(), ch
        assert pipe.meta(ch) == metas[ch], ch
        total += outs[ch].size
    assert total > 2000
    pipe.close()
