"""K2 parity: dh_demod_* (CUDA through the C ABI) vs the CPU oracle, symbol streams byte-exact.

Inputs are RRC-filtered (by the oracle) 4-level NRZ signals with clock offsets / noise so that the
variance-minimum timing loop steps (gfsk_demodulator.cpp:69-77), plus raw 2-level signals at sps 40 as in the
POCSAG pipe (examples/pocsag-decoder.sh:19-21).
"""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu


def _gpu_demod(bank, x, chunks):
    """x [C, n] numpy float32 -> list of per-channel symbol arrays, feeding the given chunk lengths."""
    C = x.shape[0]
    outs = [[] for _ in range(C)]
    pos = 0
    for c in chunks:
        d = torch.from_numpy(np.ascontiguousarray(x[:, pos:pos + c])).cuda()
        sym, nsym = bank.process(d)
        sym = sym.cpu().numpy()
        nsym = nsym.cpu().numpy()
        for ch in range(C):
            outs[ch].append(sym[ch, :nsym[ch]].copy())
        pos += c
    return [np.concatenate(o) if o else np.zeros(0, np.uint8) for o in outs]


def _signals(C, nsym, sps, seed, four_level=True, filt=True):
    orc = oracle_lib.best()
    rng = np.random.default_rng(seed)
    xs = []
    for ch in range(C):
        s = synth.random_symbols(nsym, 4 if four_level else 2, seed * 1000 + ch)
        ppm = [0, 300, -300, 800, -900, 50][ch % 6]
        snr = [None, 25, 15, 8, None, 3][ch % 6]
        x = synth.modulate(s, sps=sps, levels=synth.LEVELS4 if four_level else synth.LEVELS2, ppm=ppm,
                           phase=float(rng.integers(0, sps)), snr_db=snr, dc=[0.0, 0.05, -0.1][ch % 3], rng=rng,
                           amplitude=[0.5, 0.2, 0.9][ch % 3])
        if filt:
            x = orc.rrc(x)
        xs.append(x)
    n = min(len(x) for x in xs)
    return np.stack([x[:n] for x in xs])


@pytest.mark.parametrize("sps,four_level,invert", [(10, True, False), (20, True, False), (10, False, False),
                                                     (40, False, True), (5, True, False), (12, False, True)])
def test_demod_whole_stream(sps, four_level, invert):
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 12
    x = _signals(C, 1500, sps, seed=sps, four_level=four_level, filt=four_level)
    bank = dh.DemodBank(C, sps=sps, four_level=four_level, invert=invert)
    got = _gpu_demod(bank, x, [x.shape[1]])
    stepped = 0
    for ch in range(C):
        ref = orc.demod(x[ch], sps=sps, four_level=four_level, invert=invert)
        assert got[ch].size == ref.size, (ch, got[ch].size, ref.size)
        assert np.array_equal(got[ch], ref), "channel %d: first diff at %d" % (
            ch, int(np.argmax(got[ch] != ref)))
        stepped += ref.size != x.shape[1] // sps
    bank.close()


def test_demod_streaming_chunks():
    """Any cut of the sample stream produces the same symbol stream (carry of partial 100-symbol blocks)."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 6
    sps = 10
    x = _signals(C, 2600, sps, seed=77)
    n = x.shape[1]
    rng = np.random.default_rng(5)
    chunks = []
    left = n
    while left > 0:
        c = int(min(left, rng.choice([1, 7, 11, 12, 13, 128, 999, 1000, 1001, 1013, 2500, 4096])))
        chunks.append(c)
        left -= c
    bank = dh.DemodBank(C, sps=sps)
    got = _gpu_demod(bank, x, chunks)
    for ch in range(C):
        ref = orc.demod(x[ch], sps=sps, chunk=128)
        assert np.array_equal(got[ch], ref), ch
    bank.reset()
    got2 = _gpu_demod(bank, x[:, :5000], [5000])
    for ch in range(C):
        assert np.array_equal(got2[ch], orc.demod(x[ch, :5000], sps=sps)), ch
    bank.close()


def test_demod_degenerate_inputs():
    """All-zero, constant, negative-only (FLT_MIN max quirk, gfsk_demodulator.cpp:111) and tiny inputs."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    n = 4000
    x = np.zeros((5, n), dtype=np.float32)
    x[1] = 0.25
    x[2] = -0.5
    x[3] = np.linspace(-1, 1, n).astype(np.float32)
    x[4, ::2] = 1e-40
    bank = dh.DemodBank(5, sps=10)
    got = _gpu_demod(bank, x, [n])
    for ch in range(5):
        assert np.array_equal(got[ch], orc.demod(x[ch], sps=10)), ch
    bank.close()
    # fewer samples than one symbol needs: nothing may be emitted, everything is carried
    bank = dh.DemodBank(2, sps=10)
    y = _signals(2, 40, 10, seed=3)
    got = _gpu_demod(bank, y, [5, 6, 1, y.shape[1] - 12])
    for ch in range(2):
        assert np.array_equal(got[ch], orc.demod(y[ch], sps=10)), ch
    bank.close()


def test_demod_zero_copy_from_rrc():
    """RRC bank writing straight into the demodulator's work rows == oracle rrc -> gfsk chain."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 8
    raw = _signals(C, 3000, 10, seed=21, filt=False)
    n = raw.shape[1]
    rrc = dh.RrcBank(C, dh.RRC_WIDE)
    dem = dh.DemodBank(C, sps=10)
    chunk = 7000
    outs = [[] for _ in range(C)]
    for pos in range(0, n, chunk):
        c = min(chunk, n - pos)
        ptr, pitch = dem.reserve(chunk)      # the input rows alternate from call to call
        blk = np.zeros((C, (c + 3) & ~3), dtype=np.float32)
        blk[:, :c] = raw[:, pos:pos + c]
        d = torch.from_numpy(blk).cuda()
        import ctypes
        from digiham_b200._capi import check, lib, _stream_ptr
        check(lib().dh_rrc_process(rrc._h, d.data_ptr(), d.stride(0), ptr, pitch, c, _stream_ptr(None)))
        sym, nsym = dem.process((ptr, pitch), n=c)
        sym = sym.cpu().numpy()
        nsym = nsym.cpu().numpy()
        for ch in range(C):
            outs[ch].append(sym[ch, :nsym[ch]].copy())
    for ch in range(C):
        ref = orc.demod(orc.rrc(raw[ch]), sps=10)
        assert np.array_equal(np.concatenate(outs[ch]), ref), ch
    rrc.close()
    dem.close()
