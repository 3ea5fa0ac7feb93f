"""K6 parity: dh_dvf_* vs the CPU oracle's DigitalVoiceFilter, int16 exact (incl. clipping-range inputs)."""
import numpy as np
import pytest
import torch

import oracle_lib

pytestmark = pytest.mark.gpu


def test_dvf_bit_exact_streaming():
    import digiham_b200 as dh
    orc = oracle_lib.best()
    rng = np.random.default_rng(2)
    C, n = 70, 9000
    x = np.zeros((C, n), dtype=np.int16)
    t = np.arange(n)
    for c in range(C):
        kind = c % 5
        if kind == 0:
            x[c] = rng.integers(-32768, 32768, n)
        elif kind == 1:
            x[c] = (12000 * np.sin(2 * np.pi * (50 + 40 * c) * t / 8000)).astype(np.int16)
        elif kind == 2:
            x[c] = rng.normal(0, 3000, n).clip(-32768, 32767).astype(np.int16)
        elif kind == 3:
            x[c] = np.where((t // 37) % 2 == 0, 32767, -32768)
        else:
            x[c, ::97] = 20000
    bank = dh.DvfBank(C)
    outs = []
    pos = 0
    for c in [1, 63, 64, 65, 500, 3000, n - 3693]:
        d = torch.from_numpy(np.ascontiguousarray(x[:, pos:pos + c])).cuda()
        outs.append(bank.process(d).cpu().numpy())
        pos += c
    y = np.concatenate(outs, axis=1)
    for c in range(C):
        ref = orc.dvf(x[c], chunk=128)
        assert np.array_equal(y[c], ref), "channel %d first diff at %d" % (c, int(np.argmax(y[c] != ref)))
    bank.reset()
    d = torch.from_numpy(np.ascontiguousarray(x[:, :100])).cuda()
    assert np.array_equal(bank.process(d).cpu().numpy(), y[:, :100])
    bank.close()
