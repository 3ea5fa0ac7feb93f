"""Tiny run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck); not a test of values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C, L = 7, 9000
cases = [
    (dh.PROTO_DMR, 10, synth.LEVELS4, lambda k: synth.dmr_symbols(8, seed=k, lead_in=10)),
    (dh.PROTO_YSF, 10, synth.LEVELS4, lambda k: synth.ysf_symbols(3, seed=k, mode="mix", lead_in=10)),
    (dh.PROTO_NXDN, 20, synth.LEVELS4, lambda k: synth.nxdn_symbols(4, seed=k, lead_in=10)),
    (dh.PROTO_DSTAR, 10, synth.LEVELS2, lambda k: synth.dstar_symbols(10, seed=k, lead_in=10)),
    (dh.PROTO_POCSAG, 40, synth.LEVELS2[::-1].copy(), lambda k: synth.pocsag_bits([(5 + k, 3, "SAN")], seed=k, lead_in=3)),
]
for pid, sps, levels, gen in cases:
    nsym = L // sps + 8
    sym = np.stack([np.resize(gen(k), nsym) for k in range(C)])
    x = synth.modulate_batch(sym, L, sps=sps, levels=levels, amplitude=0.5, snr_db=12.0, seed=1, device="cuda:0")
    for mode in (False, True):
        pipe = dh.Pipe(C, pid, max_chunk=3000)
        pipe.set_async(mode)
        for pos in range(0, L, 3000):
            pipe.process(x[:, pos:pos + 3000], n=3000)
            pipe.collect()
        pipe.set_async(False)
        pipe.close()
    # symbol-level decoder with odd chunk sizes
    bank = dh.DecoderBank(C, pid)
    s8 = torch.from_numpy(sym).cuda()
    for a, b in ((0, 1), (1, 700), (700, nsym)):
        bank.process(s8[:, a:b].contiguous(), torch.full((C,), b - a, dtype=torch.int32, device="cuda"))
        bank.collect()
    bank.close()
# round 2: int16 ingest (K1 S16 variants, blocking and streaming host interface), state export / import, a world-1
# sharded pipe (wire pack kernel + compacting read-back), the device-level FEC test hooks
from digiham_b200 import shard
sym = np.stack([np.resize(synth.dmr_symbols(8, seed=k, lead_in=10), L // 10 + 8) for k in range(C)])
x = synth.modulate_batch(sym, L, sps=10, levels=synth.LEVELS4, amplitude=0.5, snr_db=12.0, seed=1, device="cuda:0")
s16 = torch.clamp(torch.round(x[:, :L] * 20000.0), -32768, 32767).to(torch.int16)
for proto, narrow in ((dh.PROTO_DMR, False), (dh.PROTO_NXDN, True)):
    pipe = dh.Pipe(C, proto, max_chunk=3000)
    for pos, c in ((0, 3000), (3000, 1), (3001, 2999), (6000, 2992)):
        blk = torch.zeros((C, (c + 7) & ~7), dtype=torch.int16, device="cuda")
        blk[:, :c] = s16[:, pos:pos + c]
        pipe.process(blk, n=c)
        pipe.collect()
    blob = pipe.export_state()
    pipe.close()
    pipe = dh.Pipe(C, proto, max_chunk=3000)
    pipe.import_state(blob)
    hb = torch.zeros((C, pipe.host_pitch_s16), dtype=torch.int16).pin_memory()
    hb[:, :3000] = s16[:, :3000].cpu()
    pipe.submit(hb, n=3000)
    pipe.submit(hb, n=3000)
    pipe.collect_step()
    pipe.collect_step()
    pipe.close()
sp = shard.ShardedPipe(C, dh.PROTO_DMR, max_chunk=3000, device="cuda:0", fmt=dh.FMT_S16)
for k in range(3):
    blk = torch.zeros((C, sp.pitch), dtype=torch.int16, device="cuda")
    blk[:, :3000] = s16[:, k * 3000:(k + 1) * 3000]
    sp.submit(blk, 3000, scatter=True)
    if k:
        sp.collect_step()
sp.collect_step()
sp.sync()
sp.close()
import ctypes
rng = np.random.default_rng(0)
for code, bits in ((0, 7), (1, 13), (2, 15), (3, 16), (4, 16), (5, 20), (6, 24), (7, 31)):
    w = rng.integers(0, 1 << bits, size=1000, dtype=np.uint32)
    ok = np.zeros(1000, dtype=np.uint8)
    dh._capi.check(dh.lib().dh_test_fec(code, w.ctypes.data, ok.ctypes.data, 1000))
pl = rng.integers(0, 256, size=(101, 25), dtype=np.uint8)
o12, okb = np.zeros((101, 12), dtype=np.uint8), np.zeros(101, dtype=np.uint8)
dh._capi.check(dh.lib().dh_test_bptc(pl.ctypes.data, o12.ctypes.data, okb.ctypes.data, 101))
for variant, steps in ((0, 100), (1, 180), (2, 36), (3, 96)):
    dd = rng.integers(0, 4, size=(33, steps), dtype=np.uint8)
    ww, mm = np.zeros((33, (steps + 31) // 32), dtype=np.uint32), np.zeros(33, dtype=np.uint32)
    dh._capi.check(dh.lib().dh_test_viterbi(variant, dd.ctypes.data, 33, ww.ctypes.data, mm.ctypes.data))
d = dh.DvfBank(C)
d.process(torch.randint(-30000, 30000, (C, 1000), dtype=torch.int16, device="cuda"))
r = dh.RrcBank(C, dh.RRC_NARROW)
r.process(torch.rand((C, 1000), device="cuda"))
torch.cuda.synchronize()
print("sanitize driver done")
