"""YSF parity: dh_decoder_* (DH_PROTO_YSF, incl. the K5 Viterbi kernel) and the rrc | gfsk | ysf pipe vs the CPU
oracle: byte stream byte-exact, metadata lines string-exact.  Streams contain header / communication / terminator
frames in V/D1, V/D2 (DN, with callsign and GPS data channels), voice-FR and data-FR modes with symbol errors, so
Viterbi + Golay(24,12) + CRC16 succeed, correct and fail."""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

MODES = ["DN", "V1", "VW", "mix", "DN", "FR"]
ERRS = [0.0, 0.004, 0.01, 0.03, 0.06]


def _streams(C, frames, seed):
    out = [synth.ysf_symbols(frames, seed=seed * 100 + ch, mode=MODES[ch % len(MODES)],
                             symbol_errors=ERRS[ch % len(ERRS)]) for ch in range(C)]
    n = min(len(s) for s in out)
    return np.stack([s[:n] for s in out])


def _check(bank, sym, chunk=0):
    orc = oracle_lib.best()
    total = 0
    for ch in range(sym.shape[0]):
        ref_out, ref_meta = orc.decode(oracle_lib.PROTO_YSF, sym[ch], chunk=chunk)
        assert bank.output(ch) == ref_out.tobytes(), "channel %d bytes differ (%d vs %d)" % (
            ch, len(bank.output(ch)), ref_out.size)
        assert bank.meta(ch) == ref_meta, "channel %d meta differs:\n%s\n---\n%s" % (
            ch, bank.meta(ch).decode(errors="replace")[:500], ref_meta.decode(errors="replace")[:500])
        total += ref_out.size + len(ref_meta)
    return total


def test_ysf_decoder_whole_stream():
    import digiham_b200 as dh
    C = 30
    sym = _streams(C, 45, seed=1)
    bank = dh.DecoderBank(C, dh.PROTO_YSF)
    bank.process(torch.from_numpy(sym).cuda(), torch.full((C,), sym.shape[1], dtype=torch.int32, device="cuda"))
    bank.collect()
    assert _check(bank, sym) > 5000
    bank.close()


def test_ysf_decoder_streaming_chunks():
    import digiham_b200 as dh
    C = 8
    sym = _streams(C, 30, seed=2)
    n = sym.shape[1]
    bank = dh.DecoderBank(C, dh.PROTO_YSF)
    rng = np.random.default_rng(3)
    pos = 0
    while pos < n:
        c = int(min(n - pos, rng.choice([1, 19, 20, 21, 479, 480, 481, 1000, 4800])))
        bank.process(torch.from_numpy(np.ascontiguousarray(sym[:, pos:pos + c])).cuda(),
                     torch.full((C,), c, dtype=torch.int32, device="cuda"))
        bank.collect()
        pos += c
    _check(bank, sym, chunk=128)
    bank.close()


def test_ysf_pipe_vs_oracle():
    """BASELINE config 3 shape (rrc_filter | gfsk_demodulator | ysf_decoder, examples/ysf-decoder.sh:19-23)."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 16
    sym = _streams(C, 25, seed=4)
    n = sym.shape[1] * 10
    x = synth.modulate_batch(sym, n, sps=10, amplitude=0.5, ppm=np.array([0, 30, -30, 60] * 4, dtype=np.float64),
                             phase=np.arange(C, dtype=np.float64) * 3, snr_db=np.array([np.inf, 20, 14, 9] * 4),
                             seed=9, device="cuda")
    pipe = dh.Pipe(C, dh.PROTO_YSF, max_chunk=40000)
    for pos in range(0, n, 40000):
        c = min(40000, n - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = x[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()
    xc = x[:, :n].cpu().numpy()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_YSF, xc, threads=8, meta_cap=1 << 15)
    total = 0
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes(), ch
        assert pipe.meta(ch) == metas[ch], ch
        total += outs[ch].size
    assert total > 2000
    pipe.close()
