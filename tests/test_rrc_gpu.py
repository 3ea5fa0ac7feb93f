"""K1 parity: dh_rrc_* (CUDA, through the C ABI) vs the CPU oracle, bit-exact.

Mirrors how the reference is exercised: one RrcFilter per channel, fed in arbitrary chunk sizes
(src/lib/cli.cpp:29-33); the reference tree has no tests of its own.
"""
import numpy as np
import pytest
import torch

import oracle_lib

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _run_gpu(bank, x, chunks):
    """x: [C, n] float32 numpy; chunks: list of chunk lengths summing to n."""
    from digiham_b200._capi import pitch4
    outs = []
    pos = 0
    for c in chunks:
        blk = np.zeros((x.shape[0], pitch4(c)), dtype=np.float32)
        blk[:, :c] = x[:, pos:pos + c]
        d = torch.from_numpy(blk).cuda()
        y = bank.process(d, n=c)
        outs.append(y[:, :c].cpu().numpy())
        pos += c
    torch.cuda.synchronize()
    return np.concatenate(outs, axis=1) if outs else np.zeros((x.shape[0], 0), np.float32)


@pytest.mark.parametrize("narrow", [False, True])
@pytest.mark.parametrize("n", [1, 3, 79, 80, 81, 161, 2175, 2176, 2177, 4353, 10000])
def test_rrc_single_call_bit_exact(narrow, n):
    import digiham_b200 as dh
    orc = oracle_lib.best()
    rng = np.random.default_rng(1000 + n)
    C = 5
    x = rng.uniform(-1, 1, size=(C, n)).astype(np.float32)
    bank = dh.RrcBank(C, dh.RRC_NARROW if narrow else dh.RRC_WIDE)
    y = _run_gpu(bank, x, [n])
    for c in range(C):
        ref = orc.rrc(x[c], narrow=narrow)
        assert np.array_equal(_bits(y[c]), _bits(ref)), "channel %d differs (max ulp %d)" % (
            c, np.abs(_bits(y[c]).astype(np.int64) - _bits(ref).astype(np.int64)).max())
    bank.close()


@pytest.mark.parametrize("narrow", [False, True])
def test_rrc_streaming_chunks_bit_exact(narrow):
    """History carry: any cut of the stream gives the same samples as one pass."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    rng = np.random.default_rng(7)
    C = 3
    chunks = [1, 2, 5, 40, 79, 81, 160, 1, 3000, 2176, 17, 4500]
    n = sum(chunks)
    x = rng.normal(0, 0.4, size=(C, n)).astype(np.float32)
    bank = dh.RrcBank(C, dh.RRC_NARROW if narrow else dh.RRC_WIDE)
    y = _run_gpu(bank, x, chunks)
    for c in range(C):
        ref = orc.rrc(x[c], narrow=narrow, chunk=128)
        assert np.array_equal(_bits(y[c]), _bits(ref))
    # reset returns to power-on state
    bank.reset()
    y2 = _run_gpu(bank, x[:, :500], [500])
    assert np.array_equal(_bits(y2), _bits(y[:, :500]))
    bank.close()


def test_rrc_special_values():
    """-0.0 products, denormals, huge values: the ordered fp32 sum must behave like the x86-64 build."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    n = 4096
    x = np.zeros((4, n), dtype=np.float32)
    x[0, ::7] = -0.0
    x[0, 100] = 1e-42
    x[1] = np.float32(1e-39)
    x[1, ::3] *= -1
    x[2, :] = np.float32(3e37)
    x[2, 50:60] = np.float32(-3.0e38)
    rng = np.random.default_rng(3)
    x[3] = (rng.integers(0, 4, n // 8).repeat(8) * 2 - 3).astype(np.float32) / 6
    bank = dh.RrcBank(4, dh.RRC_WIDE)
    y = _run_gpu(bank, x, [n])
    for c in range(4):
        ref = orc.rrc(x[c])
        assert np.array_equal(_bits(y[c]), _bits(ref)), c
    bank.close()


def test_rrc_many_channels_linearity_property():
    """Full-size property check (no oracle): filtering is shift-invariant across channels."""
    import digiham_b200 as dh
    C, n = 4096, 8704
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.rand((1, n), generator=g, device="cuda") - 0.5
    x = base.repeat(C, 1).contiguous()
    bank = dh.RrcBank(C, dh.RRC_WIDE)
    y = bank.process(x)
    torch.cuda.synchronize()
    assert torch.equal(y[0].view(torch.int32), y[C - 1].view(torch.int32))
    assert torch.equal(y.view(torch.int32), y[:1].view(torch.int32).expand_as(y))
    orc = oracle_lib.best()
    ref = orc.rrc(base[0].cpu().numpy())
    assert np.array_equal(_bits(y[17].cpu().numpy()), _bits(ref))
    bank.close()


@pytest.mark.parametrize("nz,chunks", [(32, [1000, 2000]), (160, [2999, 1]), (164, [3000]), (200, [7, 1493, 1500]),
                                       (1024, [2500, 500])])
def test_rrc_custom_taps_and_argument_errors(nz, chunks):
    """custom filters: taps in the kernel parameter block (<= 164 taps) or staged in shared memory (longer)"""
    import digiham_b200 as dh
    rng = np.random.default_rng(11)
    coeffs = rng.normal(size=nz + 1).astype(np.float32)
    gain = 3.7
    n = 3000
    x = rng.uniform(-1, 1, size=(2, n)).astype(np.float32)
    bank = dh.RrcBank(2, custom=(nz, gain, coeffs))
    y = _run_gpu(bank, x, chunks)
    # numpy restatement of src/rrc_filter/rrc_filter.cpp:22-34 for arbitrary taps
    for c in range(2):
        xp = np.concatenate([np.zeros(nz, np.float32), x[c]])
        acc = np.zeros(n, np.float32)
        for i in range(nz + 1):
            acc = (acc + (coeffs[i] * xp[i:i + n]).astype(np.float32)).astype(np.float32)
        ref = (acc.astype(np.float64) / gain).astype(np.float32)
        assert np.array_equal(_bits(y[c]), _bits(ref))
    with pytest.raises(dh.DhError):
        bank.process(torch.zeros((2, 10), device="cuda"), n=10)  # pitch not a multiple of 4
    bank.close()
    with pytest.raises(dh.DhError):
        dh.RrcBank(2, custom=(30, 1.0, np.zeros(31, np.float32)))  # nZeros not a multiple of 4


def test_rrc_more_channels_than_grid_y():
    """banks beyond 65535 channels are launched in channel slices (grid.y limit); history carries per channel"""
    import digiham_b200 as dh
    C, n = 65535 + 9, 300
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, (C, 2 * n)).astype(np.float32)
    bank = dh.RrcBank(C, dh.RRC_WIDE)
    xd = torch.from_numpy(x).cuda()
    y = torch.cat([bank.process(xd[:, :n].contiguous()), bank.process(xd[:, n:].contiguous())], dim=1).cpu().numpy()
    orc = oracle_lib.best()
    for ch in (0, 1, 65534, 65535, 65536, C - 1):
        assert np.array_equal(_bits(y[ch]), _bits(orc.rrc(x[ch]))), ch
    bank.close()
