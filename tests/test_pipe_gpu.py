"""Whole-pipe parity (BASELINE config 2 shape): RRC -> GFSK demod -> DMR decoder on the GPU (dh_pipe_*, through
the C ABI) vs the CPU oracle's rrc_filter | gfsk_demodulator | dmr_decoder chain on the same samples.

Gates: demodulated symbols byte-exact, decoder byte stream byte-exact, metadata lines
string-exact in order.  Small sizes against the oracle; the full 4096-channel size through size-independent
properties (chunking invariance, duplicate channels) plus the oracle on every distinct channel.
"""
import hashlib

import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu


def _run_pipe(x, chunk, host=False, want_symbols=False, sub_chunk=None):
    """x: float32 torch tensor [C, pitch] (cuda); returns per-channel (symbols, bytes, meta).
    sub_chunk: software-pipelining granularity (None = library default, 0 = off)."""
    import digiham_b200 as dh
    C = x.shape[0]
    n = x.shape[1]
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=chunk)
    if want_symbols:
        sub_chunk = 0        # last_symbols() only sees the final sub-chunk of a pipelined call
    if sub_chunk is not None:
        pipe.set_sub_chunk(sub_chunk)
    syms = [[] for _ in range(C)]
    for pos in range(0, n, chunk):
        c = min(chunk, n - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = x[:, pos:pos + c]
        if host:
            blk = blk.cpu().pin_memory()
        pipe.process(blk, n=c)
        pipe.collect()
        if want_symbols:
            for ch in range(C):
                syms[ch].append(pipe.last_symbols(ch))
    res = []
    for ch in range(C):
        s = np.concatenate(syms[ch]) if want_symbols else None
        res.append((s, pipe.output(ch), pipe.meta(ch)))
    pipe.close()
    return res


def test_pipe_dmr_vs_oracle():
    C, n = 96, 60000
    x, info = synth.dmr_channel_bank(C, n, seed=11, device="cuda")
    xc = x[:, :n].cpu().numpy()
    got = _run_pipe(x[:, :n], chunk=24000, want_symbols=True)
    orc = oracle_lib.best()
    syms, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, xc, threads=8, chunk=4096, want_sym=True)
    voice = 0
    for ch in range(C):
        assert np.array_equal(got[ch][0], syms[ch]), "channel %d symbols differ" % ch
        assert got[ch][1] == outs[ch].tobytes(), "channel %d bytes differ" % ch
        assert got[ch][2] == metas[ch], "channel %d meta differs" % ch
        voice += len(got[ch][1])
    assert voice > 27 * 200, "workload did not produce voice frames"
    # the software-pipelined schedule (sub-chunks on two streams) must not change a single byte
    for sub in (None, 4352, 1000):
        piped = _run_pipe(x[:, :n], chunk=24000, sub_chunk=sub)
        for ch in range(C):
            assert piped[ch][1] == got[ch][1] and piped[ch][2] == got[ch][2], (sub, ch)


def test_pipe_host_input_equals_device_input():
    C, n = 16, 30000
    x, _ = synth.dmr_channel_bank(C, n, seed=12, device="cuda")
    a = _run_pipe(x[:, :n], chunk=30000)
    b = _run_pipe(x[:, :n], chunk=7001, host=True)
    for ch in range(C):
        assert a[ch][1] == b[ch][1] and a[ch][2] == b[ch][2], ch


def test_pipe_streaming_submit_equals_blocking_calls():
    """dh_pipe_submit_host / dh_pipe_collect_step (two steps in flight, upload overlapped with decoding) produce the
    same per-channel streams as blocking process + collect calls."""
    import digiham_b200 as dh
    C, n, chunk = 24, 50000, 10000
    x, _ = synth.dmr_channel_bank(C, n, seed=14, device="cuda")
    ref = _run_pipe(x[:, :n], chunk=chunk)
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=chunk)
    blocks = []
    for pos in range(0, n, chunk):
        c = min(chunk, n - pos)
        b = torch.zeros((C, pipe.host_pitch), dtype=torch.float32).pin_memory()
        b[:, :c] = x[:, pos:pos + c].cpu()
        blocks.append((b, c))
    with pytest.raises(dh.DhError):
        pipe.collect_step()                      # nothing in flight
    pipe.submit(blocks[0][0], n=blocks[0][1])
    for b, c in blocks[1:]:
        pipe.submit(b, n=c)
        pipe.collect_step()
    with pytest.raises(dh.DhError):
        pipe.submit(blocks[0][0], n=chunk) or pipe.submit(blocks[0][0], n=chunk) or pipe.submit(blocks[0][0], n=chunk)
    pipe.close()
    # (the failed third submit above is the point of that check; run the comparison on a clean pipe)
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=chunk)
    pipe.submit(blocks[0][0], n=blocks[0][1])
    for b, c in blocks[1:]:
        pipe.submit(b, n=c)
        pipe.collect_step()
    pipe.collect_step()
    for ch in range(C):
        assert pipe.output(ch) == ref[ch][1] and pipe.meta(ch) == ref[ch][2], ch
    pipe.close()


def test_pipe_async_pipelining_equals_blocking_calls():
    """dh_pipe_set_async: K1 of call i+1 overlaps K2 + decoder of call i on internal streams; the streams carry the
    same dependencies as the single-stream order, so the results are identical (checked against the oracle)."""
    import digiham_b200 as dh
    C, n, chunk = 256, 96000, 12000
    x, _ = synth.dmr_channel_bank(C, n, seed=17, device="cuda")
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=chunk)
    pipe.set_async(True)
    for k, pos in enumerate(range(0, n, chunk)):
        pipe.process(x[:, pos:pos + chunk], n=chunk)
        if k % 2 == 1:
            pipe.collect()           # syncs, reads back (result buffers hold two calls)
    pipe.collect()
    pipe.set_async(False)
    orc = oracle_lib.best()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, x[:, :n].cpu().numpy(), threads=8)
    total = 0
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes() and pipe.meta(ch) == metas[ch], ch
        total += len(outs[ch])
    assert total > 27 * 200
    pipe.close()


def test_pipe_async_host_input():
    """asynchronous mode with HOST input: the staging copy of call i+1 must wait for K1 of call i"""
    import digiham_b200 as dh
    C, n, chunk = 64, 60000, 6000
    x, _ = synth.dmr_channel_bank(C, n, seed=19, device="cpu")
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=chunk)
    pipe.set_async(True)
    blocks = []
    for pos in range(0, n, chunk):
        b = torch.zeros((C, pipe.host_pitch), dtype=torch.float32).pin_memory()
        b[:, :chunk] = x[:, pos:pos + chunk]
        blocks.append(b)
    for k, b in enumerate(blocks):
        pipe.process(b, n=chunk)
        if k % 2 == 1:
            pipe.collect()
    pipe.collect()
    _, outs, metas = oracle_lib.best().pipe_batch(oracle_lib.PROTO_DMR, x[:, :n].numpy(), threads=8)
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes() and pipe.meta(ch) == metas[ch], ch
    pipe.close()


def test_pipe_full_size_properties():
    """4096 channels (BASELINE config 2): results must not depend on the chunking, and duplicated channels must
    produce identical streams (no cross-channel interference)."""
    C, n = 4096, 48000
    x, _ = synth.dmr_channel_bank(C, n, seed=13, device="cuda")
    x[C // 2:] = x[:C // 2]          # second half duplicates the first
    a = _run_pipe(x[:, :n], chunk=48000)
    b = _run_pipe(x[:, :n], chunk=16384)

    def digest(res):
        h = hashlib.sha256()
        for s, o, m in res:
            h.update(hashlib.sha256(o).digest())
            h.update(hashlib.sha256(m).digest())
        return h.hexdigest()

    assert digest(a) == digest(b)
    for ch in range(C // 2):
        assert a[ch][1] == a[ch + C // 2][1] and a[ch][2] == a[ch + C // 2][2]
    assert sum(len(o) for _, o, _ in a) > 27 * 5000
    # every distinct channel of the full-size run against the oracle (all host cores)
    import os
    orc = oracle_lib.best()
    xc = x[:C // 2, :n].cpu().numpy()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, xc, threads=os.cpu_count() or 8, meta_cap=1 << 15)
    for ch in range(C // 2):
        assert a[ch][1] == outs[ch].tobytes() and a[ch][2] == metas[ch], ch


def test_pipe_config1_ten_seconds_single_call():
    """BASELINE configs[0] shape: one channel, 10 s (480000 samples) in ONE call, plus two more channels beside it."""
    import digiham_b200 as dh
    C, n = 3, 480000
    x, _ = synth.dmr_channel_bank(C, n, seed=23, device="cuda", noise_fraction=0.0)
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=n)
    pipe.process(x, n=n)
    pipe.collect()
    _, outs, metas = oracle_lib.best().pipe_batch(oracle_lib.PROTO_DMR, x[:, :n].cpu().numpy(), threads=3, meta_cap=1 << 16)
    for ch in range(C):
        assert pipe.output(ch) == outs[ch].tobytes() and pipe.meta(ch) == metas[ch], ch
    assert sum(len(o) for o in outs) > 27 * 100
    pipe.close()
