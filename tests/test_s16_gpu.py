"""int16 ingest: `csdr convert -i s16 -o float` fused into K1.

The reference pipe starts at s16 (reference examples/dmr-decoder.sh:13-15: rtl_fm | csdr convert -i s16 -o float |
csdr dcblock | rrc_filter ...).  csdr's source is not part of the reference tree; its s16 -> float conversion is
`(float) in / SHRT_MAX` (one IEEE float32 division), restated here as numpy float32(s) / float32(32767) feeding the
compiled reference modules — the gate is bit-exactness of everything downstream of that conversion.
"""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu


def _to_s16(x):
    """float32 torch [C, n] in about [-1, 1] -> int16 like an rtl_fm discriminator output."""
    return torch.clamp(torch.round(x * 20000.0), -32768, 32767).to(torch.int16)


def _s16_to_f32_numpy(s):
    return s.astype(np.float32) / np.float32(32767)


def test_rrc_s16_all_values_and_chunking():
    import digiham_b200 as dh
    # every int16 value appears (conversion exactness on the device), then random data; ragged chunks incl. tiny ones
    rng = np.random.default_rng(5)
    C, n = 6, 70000
    s = rng.integers(-32768, 32768, size=(C, n)).astype(np.int16)
    s[0, :65536] = np.arange(-32768, 32768, dtype=np.int16)
    s[1, :65536] = np.arange(32767, -32769, -1, dtype=np.int16)
    orc = oracle_lib.best()
    for narrow in (False, True):
        want = np.stack([orc.rrc(_s16_to_f32_numpy(s[c]), narrow=narrow) for c in range(C)])
        bank = dh.RrcBank(C, dh.RRC_NARROW if narrow else dh.RRC_WIDE)
        got = np.empty((C, n), dtype=np.float32)
        pos = 0
        for chunk in (1, 7, 40, 2432, 19999, 31, 48000):
            c = min(chunk, n - pos)
            if c <= 0:
                break
            blk = torch.zeros((C, (c + 7) & ~7), dtype=torch.int16, device="cuda")
            blk[:, :c] = torch.from_numpy(s[:, pos:pos + c]).cuda()
            got[:, pos:pos + c] = bank.process(blk, n=c)[:, :c].cpu().numpy()
            pos += c
        assert pos == n
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "narrow=%s" % narrow
        bank.close()


def test_rrc_mixed_float_and_s16_calls_share_one_history():
    import digiham_b200 as dh
    rng = np.random.default_rng(6)
    C, n = 3, 9000
    s = rng.integers(-30000, 30000, size=(C, n)).astype(np.int16)
    f = _s16_to_f32_numpy(s)
    orc = oracle_lib.best()
    want = np.stack([orc.rrc(f[c]) for c in range(C)])
    bank = dh.RrcBank(C)
    got = np.empty_like(want)
    half = 4000
    a = torch.zeros((C, half), dtype=torch.float32, device="cuda")
    a[:] = torch.from_numpy(f[:, :half]).cuda()
    got[:, :half] = bank.process(a, n=half)[:, :half].cpu().numpy()
    rest = n - half
    b = torch.zeros((C, (rest + 7) & ~7), dtype=torch.int16, device="cuda")
    b[:, :rest] = torch.from_numpy(s[:, half:]).cuda()
    got[:, half:] = bank.process(b, n=rest)[:, :rest].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("proto_name", ["dmr", "nxdn"])
def test_pipe_s16_device_and_host_vs_oracle(proto_name):
    import digiham_b200 as dh
    C, n = 48, 48000
    if proto_name == "dmr":
        x, _ = synth.dmr_channel_bank(C, n, seed=31, device="cuda")
        dh_proto, orc_proto = dh.PROTO_DMR, oracle_lib.PROTO_DMR
    else:
        sym = np.stack([np.resize(synth.nxdn_symbols(16, seed=700 + k, lead_in=0), n // 20 + 8) for k in range(C)])
        x = synth.modulate_batch(sym, n, sps=20, levels=synth.LEVELS4, amplitude=np.full(C, 0.5), ppm=np.zeros(C),
                                 phase=np.zeros(C), snr_db=np.full(C, 20.0), seed=3, device="cuda")
        dh_proto, orc_proto = dh.PROTO_NXDN, oracle_lib.PROTO_NXDN
    s = _to_s16(x[:, :n])
    f = _s16_to_f32_numpy(s.cpu().numpy())
    orc = oracle_lib.best()
    syms, outs, metas = orc.pipe_batch(orc_proto, f, threads=8, chunk=4096, want_sym=True)
    assert sum(len(o) for o in outs) > 0, "workload produced no frames"

    # device blocks, ragged chunks
    pipe = dh.Pipe(C, dh_proto, max_chunk=20000)
    pos = 0
    got_sym = [[] for _ in range(C)]
    for chunk in (20000, 12345, 8, 15647):
        c = min(chunk, n - pos)
        blk = torch.zeros((C, (c + 7) & ~7), dtype=torch.int16, device="cuda")
        blk[:, :c] = s[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()
        for ch in range(0, C, 7):
            got_sym[ch].append(pipe.last_symbols(ch))
        pos += c
    assert pos == n
    for ch in range(C):
        if ch % 7 == 0:
            assert np.array_equal(np.concatenate(got_sym[ch]), syms[ch]), ch
        assert pipe.output(ch) == outs[ch].tobytes(), ch
        assert pipe.meta(ch) == metas[ch], ch
    pipe.close()

    # streaming host interface with pinned int16 blocks (what the e2e_s16 bench arm does)
    pipe = dh.Pipe(C, dh_proto, max_chunk=12000)
    pitch = pipe.host_pitch_s16
    bufs = []
    for pos in range(0, n, 12000):
        b = torch.zeros((C, pitch), dtype=torch.int16).pin_memory()
        b[:, :12000] = s[:, pos:pos + 12000].cpu()
        bufs.append(b)
    pipe.submit(bufs[0], n=12000)
    for i in range(1, len(bufs)):
        pipe.submit(bufs[i], n=12000)
        pipe.collect_step()
    pipe.collect_step()
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes(), ch
        assert pipe.meta(ch) == metas[ch], ch
    pipe.close()


def test_s16_is_refused_by_pipes_without_rrc_stage():
    import digiham_b200 as dh
    pipe = dh.Pipe(4, dh.PROTO_POCSAG, max_chunk=4000)
    blk = torch.zeros((4, 4000), dtype=torch.int16, device="cuda")
    with pytest.raises(dh.DhError):
        pipe.process(blk, n=4000)
    pipe.close()
