"""K2 split schedule (dh_demod_set_split: search chain -> per-symbol window sums -> per-block slicing, three kernels)
against the compiled reference and against the one-kernel schedule: symbol streams byte-exact for every sps variant,
any chunking (partial 100-symbol blocks carried across calls), caller rows read in place and work rows, schedule
changes in mid-stream, and the whole DMR / NXDN / D-Star pipes in blocking and asynchronous mode.
"""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth
from test_demod_gpu import _gpu_demod, _signals

pytestmark = pytest.mark.gpu


def _chunks(n, rng, sps):
    menu = [1, 2, 7, sps, sps + 1, sps + 2, 3 * sps, 99 * sps, 100 * sps - 1, 100 * sps, 100 * sps + 1, 101 * sps,
            250 * sps, 37 * sps + 3, 1013, 4096]
    out = []
    left = n
    while left > 0:
        c = int(min(left, rng.choice(menu)))
        out.append(c)
        left -= c
    return out


@pytest.mark.parametrize("sps,four_level,invert", [(10, True, False), (20, True, False), (10, False, False),
                                                     (40, False, True), (5, True, False), (12, False, True),
                                                     (25, True, False)])
def test_split_equals_reference(sps, four_level, invert):
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C = 13
    x = _signals(C, 2300, sps, seed=100 + sps, four_level=four_level, filt=four_level)
    n = x.shape[1]
    ref = [orc.demod(x[ch], sps=sps, four_level=four_level, invert=invert) for ch in range(C)]
    rng = np.random.default_rng(sps)
    for trial, chunks in enumerate([[n], _chunks(n, rng, sps), _chunks(n, rng, sps)]):
        bank = dh.DemodBank(C, sps=sps, four_level=four_level, invert=invert)
        bank.set_split(True)
        assert bank.kernels_per_call == 3
        got = _gpu_demod(bank, x, chunks)
        for ch in range(C):
            assert got[ch].size == ref[ch].size, (trial, ch, got[ch].size, ref[ch].size)
            assert np.array_equal(got[ch], ref[ch]), "trial %d channel %d: first diff at %d" % (
                trial, ch, int(np.argmax(got[ch] != ref[ch])))
        # the bank keeps working after a reset
        bank.reset()
        got = _gpu_demod(bank, x[:, :n // 3], [n // 3])
        for ch in range(C):
            assert np.array_equal(got[ch], orc.demod(x[ch, :n // 3], sps=sps, four_level=four_level, invert=invert)), ch
        bank.close()


def test_split_toggled_in_mid_stream_and_unaligned_rows():
    """The schedule may change between any two calls; rows that are not 16-byte aligned take the copy path."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C, sps = 7, 10
    x = _signals(C, 3100, sps, seed=9)
    n = x.shape[1]
    rng = np.random.default_rng(3)
    chunks = _chunks(n, rng, sps)
    bank = dh.DemodBank(C, sps=sps)
    outs = [[] for _ in range(C)]
    pos = 0
    for k, c in enumerate(chunks):
        bank.set_split(k % 3 != 1)
        # every other block starts one float off a 16-byte boundary
        buf = torch.zeros((C, c + 8), dtype=torch.float32, device="cuda")
        off = 1 if k % 2 else 0
        view = buf[:, off:off + c]
        view.copy_(torch.from_numpy(np.ascontiguousarray(x[:, pos:pos + c])))
        sym, nsym = bank.process(view, n=c)
        sym = sym.cpu().numpy()
        nsym = nsym.cpu().numpy()
        for ch in range(C):
            outs[ch].append(sym[ch, :nsym[ch]].copy())
        pos += c
    for ch in range(C):
        assert np.array_equal(np.concatenate(outs[ch]), orc.demod(x[ch], sps=sps)), ch
    bank.close()


def test_split_degenerate_inputs():
    import digiham_b200 as dh
    orc = oracle_lib.best()
    n = 4000
    x = np.zeros((5, n), dtype=np.float32)
    x[1] = 0.25
    x[2] = -0.5
    x[3] = np.linspace(-1, 1, n).astype(np.float32)
    x[4, ::2] = 1e-40
    bank = dh.DemodBank(5, sps=10)
    bank.set_split(True)
    got = _gpu_demod(bank, x, [1500, 1, 2499])
    for ch in range(5):
        assert np.array_equal(got[ch], orc.demod(x[ch], sps=10)), ch
    bank.close()
    bank = dh.DemodBank(2, sps=10)
    bank.set_split(True)
    y = _signals(2, 40, 10, seed=3)
    got = _gpu_demod(bank, y, [5, 6, 1, y.shape[1] - 12])
    for ch in range(2):
        assert np.array_equal(got[ch], orc.demod(y[ch], sps=10)), ch
    bank.close()


def _run(proto, x, n, chunk, split, async_mode):
    import digiham_b200 as dh
    C = x.shape[0]
    pipe = dh.Pipe(C, proto, max_chunk=chunk)
    pipe.set_demod_split(split)
    assert pipe.demod_kernels_per_call == (3 if split else 1)
    if async_mode:
        pipe.set_async(True)
    for pos in range(0, n, chunk):
        c = min(chunk, n - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = x[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()           # syncs an asynchronous pipe first; the device result slots hold two calls at most
    res = [(pipe.output(ch), pipe.meta(ch)) for ch in range(C)]
    pipe.close()
    return res


def test_split_pipe_dmr_vs_reference_blocking_and_async():
    import digiham_b200 as dh
    C, n = 192, 72000
    x, _ = synth.dmr_channel_bank(C, n, seed=31, device="cuda")
    orc = oracle_lib.best()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, x[:, :n].cpu().numpy(), threads=8, chunk=4096)
    total = 0
    for chunk, async_mode in ((24000, False), (9000, True), (7001, False)):
        got = _run(dh.PROTO_DMR, x, n, chunk, True, async_mode)
        for ch in range(C):
            assert got[ch][0] == outs[ch].tobytes(), (chunk, async_mode, ch)
            assert got[ch][1] == metas[ch], (chunk, async_mode, ch)
            total += len(got[ch][0])
    assert total > 27 * 500


def _other_pipe_input(name, C, n):
    from test_full_size_gpu import _bank_signal
    if name == "ysf":
        pool = np.stack([synth.ysf_symbols(8, seed=900 + k, mode=["DN", "V1", "VW", "mix"][k % 4], lead_in=0)[:2880]
                         for k in range(8)])
        return _bank_signal(pool, C, n, 10, synth.LEVELS4, seed=41)
    if name == "pocsag":
        texts = ["HELLO B200", "THE QUICK BROWN FOX", "73", "x" * 40]
        pool = []
        for k in range(8):
            bits = synth.pocsag_bits([(1000 + k, 3, texts[k % 4]), (77 + k, 3, texts[(k + 1) % 4])], seed=k,
                                     bit_errors=k % 3, lead_in=0, preamble=200, trailing_batches=1)
            pool.append(np.resize(bits, 1200))
        return _bank_signal(np.stack(pool), C, n, 40, synth.LEVELS2[::-1].copy(), seed=42)
    if name == "nxdn":
        pool = np.stack([synth.nxdn_symbols(16, seed=950 + k, lead_in=0)[:2400] for k in range(8)])
        return _bank_signal(pool, C, n, 20, synth.LEVELS4, seed=43)
    pool = np.stack([np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), 100),
                                     synth.dstar_symbols(60, seed=970 + k, lead_in=0)])[:4800] for k in range(8)])
    return _bank_signal(pool, C, n, 10, synth.LEVELS2, seed=44)


@pytest.mark.parametrize("proto_name", ["nxdn", "dstar", "pocsag", "ysf"])
def test_split_other_pipes_equal_reference(proto_name):
    """Narrow RRC + sps 20 (NXDN), caller rows read in place at sps 10 / 40 (D-Star, POCSAG), YSF."""
    import digiham_b200 as dh
    proto = {"nxdn": dh.PROTO_NXDN, "dstar": dh.PROTO_DSTAR, "pocsag": dh.PROTO_POCSAG, "ysf": dh.PROTO_YSF}[proto_name]
    orc_proto = {"nxdn": oracle_lib.PROTO_NXDN, "dstar": oracle_lib.PROTO_DSTAR, "pocsag": oracle_lib.PROTO_POCSAG,
                 "ysf": oracle_lib.PROTO_YSF}[proto_name]
    C, n = 64, (24000 if proto_name == "ysf" else 48000)
    x = _other_pipe_input(proto_name, C, n)
    got = _run(proto, x, n, 11003, True, False)
    orc = oracle_lib.best()
    _, outs, metas = orc.pipe_batch(orc_proto, x[:, :n].cpu().numpy(), threads=8, meta_cap=1 << 15)
    assert sum(len(r[0]) + len(r[1]) for r in got) > 0
    for ch in range(C):
        assert got[ch][0] == outs[ch].tobytes() and got[ch][1] == metas[ch], ch


def test_split_auto_policy_follows_bank_and_call_size():
    """Default (-1): small banks use the three-kernel schedule for calls that span many blocks, the one-kernel
    schedule for short calls and for large banks; the symbol stream does not notice the changes."""
    import digiham_b200 as dh
    orc = oracle_lib.best()
    C, sps = 5, 10
    x = _signals(C, 2600, sps, seed=17)
    n = x.shape[1]
    bank = dh.DemodBank(C, sps=sps)
    bank.set_split(None)
    outs = [[] for _ in range(C)]
    pos = 0
    seen = set()
    for c in [12000, 700, 3000, 9000, n - 24700]:
        d = torch.from_numpy(np.ascontiguousarray(x[:, pos:pos + c])).cuda()
        sym, nsym = bank.process(d)
        seen.add((c >= 8000, bank.kernels_per_call))
        sym, nsym = sym.cpu().numpy(), nsym.cpu().numpy()
        for ch in range(C):
            outs[ch].append(sym[ch, :nsym[ch]].copy())
        pos += c
    assert seen == {(True, 3), (False, 1)}, seen
    for ch in range(C):
        assert np.array_equal(np.concatenate(outs[ch]), orc.demod(x[ch], sps=sps)), ch
    bank.close()
    big = dh.DemodBank(2048, sps=sps)
    big.process(torch.zeros((2048, 12000), dtype=torch.float32, device="cuda"))
    assert big.kernels_per_call == 1
    big.close()
