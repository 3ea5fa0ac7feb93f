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
d = dh.DvfBank(C)
d.process(torch.randint(-30000, 30000, (C, 1000), dtype=torch.int16, device="cuda"))
r = dh.RrcBank(C, dh.RRC_NARROW)
r.process(torch.rand((C, 1000), device="cuda"))
torch.cuda.synchronize()
print("sanitize driver done")
