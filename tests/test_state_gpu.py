"""Per-channel state export / import (dh_*_state_*): a bank that is exported, destroyed, re-created and re-loaded must
continue every channel's stream byte-identically — checked against an uninterrupted run and, for the DMR pipe,
against the compiled reference fed with the whole stream."""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu


def _signal(name, C, n):
    import bench
    return bench.workload_signal(name, C, n, 77, "cuda")[:, :n].contiguous()


def _feed(pipe, x, lo, hi, chunk):
    C = x.shape[0]
    for pos in range(lo, hi, chunk):
        c = min(chunk, hi - pos)
        blk = torch.zeros((C, (c + 3) & ~3), dtype=torch.float32, device="cuda")
        blk[:, :c] = x[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()


@pytest.mark.parametrize("name", ["dmr", "ysf", "nxdn", "dstar", "pocsag"])
@pytest.mark.parametrize("async_mode", [False, True])
def test_pipe_state_roundtrip_continues_stream(name, async_mode):
    import bench
    import digiham_b200 as dh
    C, n, cut, chunk = 24, 72000, 33333, 12000
    x = _signal(name, C, n)
    dh_proto, orc_proto = bench.proto_ids(name)
    whole = dh.Pipe(C, dh_proto, max_chunk=chunk)
    _feed(whole, x, 0, n, chunk)
    want = [(whole.output(c), whole.meta(c)) for c in range(C)]
    whole.close()
    assert sum(len(o) + len(m) for o, m in want) > 0

    a = dh.Pipe(C, dh_proto, max_chunk=chunk)
    a.set_async(async_mode)
    _feed(a, x, 0, cut, chunk)
    first = [(a.output(c), a.meta(c)) for c in range(C)]
    blob = a.export_state()
    a.close()
    del a
    b = dh.Pipe(C, dh_proto, max_chunk=chunk)
    b.set_async(async_mode)
    b.import_state(blob)
    _feed(b, x, cut, n, chunk)
    for c in range(C):
        assert first[c][0] + b.output(c) == want[c][0], (name, c)
        assert first[c][1] + b.meta(c) == want[c][1], (name, c)
    # a blob of one configuration is refused by any other
    other = dh.Pipe(C + 1, dh_proto, max_chunk=chunk)
    with pytest.raises(dh.DhError):
        other.import_state(blob)
    other.close()
    with pytest.raises(dh.DhError):
        b.import_state(blob[:len(blob) // 2])
    b.close()
    if name == "dmr":
        _, outs, metas = oracle_lib.best().pipe_batch(orc_proto, x.cpu().numpy(), threads=8, meta_cap=1 << 15)
        for c in range(C):
            assert want[c][0] == outs[c].tobytes() and want[c][1] == metas[c], c


def test_rrc_and_demod_bank_state_roundtrip():
    import digiham_b200 as dh
    C, n, cut = 5, 30000, 12345
    x = _signal("dmr", C, n)
    orc = oracle_lib.best()
    xc = x.cpu().numpy()
    want_f = np.stack([orc.rrc(xc[c]) for c in range(C)])
    want_s = [orc.demod(want_f[c], sps=10) for c in range(C)]

    def blk(t, lo, hi):
        b = torch.zeros((C, (hi - lo + 3) & ~3), dtype=torch.float32, device="cuda")
        b[:, :hi - lo] = t[:, lo:hi]
        return b

    r1 = dh.RrcBank(C)
    f1 = r1.process(blk(x, 0, cut), n=cut)[:, :cut]
    blob = r1.export_state()
    r1.close()
    r2 = dh.RrcBank(C)
    r2.import_state(blob)
    f2 = r2.process(blk(x, cut, n), n=n - cut)[:, :n - cut]
    filt = torch.cat([f1, f2], dim=1)
    assert np.array_equal(filt.cpu().numpy().view(np.uint32), want_f.view(np.uint32))
    narrow = dh.RrcBank(C, dh.RRC_NARROW)
    with pytest.raises(dh.DhError):
        narrow.import_state(blob)

    d1 = dh.DemodBank(C, sps=10)
    s1, n1 = d1.process(blk(filt, 0, cut), n=cut)
    blob = d1.export_state()
    d1.close()
    d2 = dh.DemodBank(C, sps=10)
    d2.import_state(blob)
    s2, n2 = d2.process(blk(filt, cut, n), n=n - cut)
    for c in range(C):
        got = np.concatenate([s1[c, :int(n1[c])].cpu().numpy(), s2[c, :int(n2[c])].cpu().numpy()])
        assert np.array_equal(got, want_s[c]), c
    d3 = dh.DemodBank(C, sps=20)
    with pytest.raises(dh.DhError):
        d3.import_state(blob)
