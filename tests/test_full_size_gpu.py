"""Full-size runs of the BASELINE.json configurations that are not the bench line: 8192-channel YSF pipe
(configs[2]) and 32768-channel POCSAG pipe (configs[4]), plus many-channel NXDN / D-Star pipes.  EVERY distinct
channel is compared byte-exactly with the oracle (all host cores, a few thousand channels per slab); the second half
of each bank duplicates the first and must reproduce it channel by channel (no cross-channel interference), and
the results must not depend on how the stream is cut into process calls."""
import os
import hashlib

import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu


def _bank_signal(symbols, channels, n, sps, levels, seed):
    """tile a pool of symbol streams over `channels` channels with per-channel shift / impairments"""
    rng = np.random.default_rng(seed)
    pool, S = symbols.shape
    sym = np.empty((channels, S), dtype=np.uint8)
    shifts = rng.integers(0, S, size=channels)
    for c in range(channels):
        sym[c] = np.roll(symbols[c % pool], int(shifts[c]))
    snr = rng.choice([np.inf, 20.0, 12.0, 8.0], size=channels)
    ppm = rng.choice([0.0, 20.0, -20.0, 50.0, -50.0], size=channels)
    phase = rng.integers(0, 4 * sps, size=channels).astype(np.float64)
    amp = rng.choice([0.25, 0.5, 0.8], size=channels)
    x = synth.modulate_batch(sym, n, sps=sps, levels=levels, amplitude=amp, ppm=ppm, phase=phase, snr_db=snr,
                             seed=seed + 1, device="cuda")
    half = channels // 2
    x[half:] = x[:half]              # second half duplicates the first
    return x


def _run(proto, x, n, chunk):
    import digiham_b200 as dh
    C = x.shape[0]
    pipe = dh.Pipe(C, proto, max_chunk=chunk)
    for pos in range(0, n, chunk):
        c = min(chunk, n - pos)
        if pos == 0 and c == n and x.shape[1] % 4 == 0:
            pipe.process(x, n=c)
        else:
            blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
            blk[:, :c] = x[:, pos:pos + c]
            pipe.process(blk, n=c)
        pipe.collect()
    res = [(pipe.output(ch), pipe.meta(ch)) for ch in range(C)]
    pipe.close()
    return res


def _digest(res):
    h = hashlib.sha256()
    for o, m in res:
        h.update(hashlib.sha256(o).digest())
        h.update(hashlib.sha256(m).digest())
    return h.hexdigest()


def _full_size(proto, orc_proto, x, n, chunk_b, min_bytes, slab=2048):
    C = x.shape[0]
    a = _run(proto, x, n, chunk=n)
    b = _run(proto, x, n, chunk=chunk_b)
    assert _digest(a) == _digest(b)
    half = C // 2
    for ch in range(half):
        assert a[ch] == a[ch + half], ch
    assert sum(len(o) + len(m) for o, m in a) > min_bytes
    orc = oracle_lib.best()
    threads = os.cpu_count() or 8
    for c0 in range(0, half, slab):
        c1 = min(half, c0 + slab)
        xc = x[c0:c1, :n].cpu().numpy()
        _, outs, metas = orc.pipe_batch(orc_proto, xc, threads=threads, meta_cap=1 << 15)
        for ch in range(c0, c1):
            assert a[ch][0] == outs[ch - c0].tobytes() and a[ch][1] == metas[ch - c0], ch


def test_ysf_pipe_8192_channels():
    """BASELINE configs[2]: 8192 ch RRC -> GFSK demod -> ysf_decoder (Viterbi FEC)."""
    import digiham_b200 as dh
    C, n = 8192, 24000
    pool = np.stack([synth.ysf_symbols(8, seed=900 + k, mode=["DN", "V1", "VW", "mix"][k % 4], lead_in=0)[:2880]
                     for k in range(24)])
    x = _bank_signal(pool, C, n, 10, synth.LEVELS4, seed=21)
    _full_size(dh.PROTO_YSF, oracle_lib.PROTO_YSF, x, n, 10000, 40 * C)


def test_pocsag_pipe_32768_channels():
    """BASELINE configs[4]: 32768 ch FskDemodulator(40, invert) -> pocsag_decoder, 1200 bit/s."""
    import digiham_b200 as dh
    C, n = 32768, 48000
    texts = ["HELLO B200", "THE QUICK BROWN FOX", "73", "x" * 40]
    pool = []
    for k in range(16):
        bits = synth.pocsag_bits([(1000 + k, 3, texts[k % 4]), (77 + k, 3, texts[(k + 1) % 4])], seed=k,
                                 bit_errors=k % 3, lead_in=0, preamble=200, trailing_batches=1)
        pool.append(np.resize(bits, 1200))
    x = _bank_signal(np.stack(pool), C, n, 40, synth.LEVELS2[::-1].copy(), seed=22)
    _full_size(dh.PROTO_POCSAG, oracle_lib.PROTO_POCSAG, x, n, 20000, 4 * C)


def test_nxdn_pipe_4096_channels():
    import digiham_b200 as dh
    C, n = 4096, 48000
    pool = np.stack([synth.nxdn_symbols(16, seed=950 + k, lead_in=0)[:2400] for k in range(24)])
    x = _bank_signal(pool, C, n, 20, synth.LEVELS4, seed=23)
    _full_size(dh.PROTO_NXDN, oracle_lib.PROTO_NXDN, x, n, 20000, 30 * C)


def test_dstar_pipe_4096_channels():
    import digiham_b200 as dh
    C, n = 4096, 48000
    pool = np.stack([np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), 100),
                                     synth.dstar_symbols(60, seed=970 + k, lead_in=0)])[:4800] for k in range(24)])
    x = _bank_signal(pool, C, n, 10, synth.LEVELS2, seed=24)
    _full_size(dh.PROTO_DSTAR, oracle_lib.PROTO_DSTAR, x, n, 20000, 20 * C)
