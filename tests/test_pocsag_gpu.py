"""POCSAG parity: dh_decoder_* (DH_PROTO_POCSAG) and the FskDemodulator(40, invert) -> decoder pipe vs the CPU
oracle; output text byte-exact.  Streams carry alphanumeric messages with 0..3 bit errors per codeword so that BCH
correction, BCH failure (message dropped) and sync loss all occur."""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

TEXTS = ["HELLO B200", "THE QUICK BROWN FOX JUMPS OVER THE LAZY DOG 0123456789", "A", "x" * 90, "73 DE DL1ABC",
         "[]{}~|\\^_`@#$%&*()", ""]


def _streams(C, seed):
    rng = np.random.default_rng(seed)
    out = []
    for ch in range(C):
        msgs = []
        for k in range(int(rng.integers(1, 6))):
            fn = int(rng.choice([3, 3, 3, 1, 0, 2]))
            msgs.append((int(rng.integers(8, 1 << 21)), fn, TEXTS[int(rng.integers(0, len(TEXTS)))]))
        bits = synth.pocsag_bits(msgs, seed=seed * 100 + ch, bit_errors=ch % 4, lead_in=int(rng.integers(0, 100)),
                                 trailing_batches=2)
        # some channels lose the carrier in the middle: noise, then a second transmission
        if ch % 5 == 4:
            bits = np.concatenate([bits, rng.integers(0, 2, 700).astype(np.uint8),
                                   synth.pocsag_bits(msgs[:1], seed=ch, trailing_batches=1)])
        out.append(bits)
    n = max(len(b) for b in out)
    res = np.zeros((C, n), dtype=np.uint8)
    lens = []
    for ch, b in enumerate(out):
        res[ch, :len(b)] = b
        lens.append(len(b))
    return res, lens


def test_pocsag_decoder_vs_oracle():
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 24
    bits, lens = _streams(C, seed=3)
    bank = dh.DecoderBank(C, dh.PROTO_POCSAG)
    bank.process(torch.from_numpy(bits).cuda(), torch.tensor(lens, dtype=torch.int32, device="cuda"))
    bank.collect()
    total = 0
    for ch in range(C):
        ref, _ = orc.decode(oracle_lib.PROTO_POCSAG, bits[ch, :lens[ch]])
        assert bank.output(ch) == ref.tobytes(), "channel %d:\n%r\n%r" % (ch, bank.output(ch), ref.tobytes())
        assert bank.meta(ch) == b""
        total += ref.size
    assert total > 500
    bank.close()


def test_pocsag_decoder_streaming_chunks():
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 8
    bits, lens = _streams(C, seed=4)
    n = min(lens)
    bank = dh.DecoderBank(C, dh.PROTO_POCSAG)
    rng = np.random.default_rng(1)
    pos = 0
    while pos < n:
        c = int(min(n - pos, rng.choice([1, 31, 32, 33, 64, 100, 544, 1000])))
        bank.process(torch.from_numpy(np.ascontiguousarray(bits[:, pos:pos + c])).cuda(),
                     torch.full((C,), c, dtype=torch.int32, device="cuda"))
        bank.collect()
        pos += c
    for ch in range(C):
        ref, _ = orc.decode(oracle_lib.PROTO_POCSAG, bits[ch, :n], chunk=128)
        assert bank.output(ch) == ref.tobytes(), ch
    bank.close()


def test_pocsag_pipe_vs_oracle():
    """fsk_demodulator -i -s 40 | pocsag_decoder (examples/pocsag-decoder.sh:19-21) on noisy 2-level signals."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 12
    bits, lens = _streams(C, seed=5)
    n_bits = min(lens)
    rng = np.random.default_rng(7)
    xs = []
    for ch in range(C):
        xs.append(synth.modulate(bits[ch, :n_bits], sps=40, levels=synth.LEVELS2[::-1].copy(),
                                 ppm=[0, 150, -150][ch % 3], phase=float(rng.integers(0, 40)),
                                 snr_db=[None, 20, 12][ch % 3], rng=rng, amplitude=[0.5, 0.3][ch % 2]))
    n = min(len(x) for x in xs)
    x = np.stack([v[:n] for v in xs])
    pipe = dh.Pipe(C, dh.PROTO_POCSAG, max_chunk=50000)
    for pos in range(0, n, 50000):
        c = min(50000, n - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = torch.from_numpy(x[:, pos:pos + c]).cuda()
        pipe.process(blk, n=c)
        pipe.collect()
    total = 0
    for ch in range(C):
        _, ref, _ = orc.pipe(oracle_lib.PROTO_POCSAG, x[ch])
        assert pipe.output(ch) == ref.tobytes(), ch
        total += ref.size
    assert total > 100
    pipe.close()
