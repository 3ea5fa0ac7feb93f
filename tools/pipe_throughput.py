"""Device-resident throughput of every pipe (tuning / documentation aid, not the bench contract).
usage: pipe_throughput.py [steps]   -> one line per BASELINE-style configuration"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import digiham_b200 as dh
from digiham_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = 48000
CASES = [
    ("dmr", 4096, 10, synth.LEVELS4, dh.PROTO_DMR, None),
    ("ysf", 8192, 10, synth.LEVELS4, dh.PROTO_YSF, lambda k: synth.ysf_symbols(12, seed=k, mode="mix", lead_in=0)),
    ("nxdn", 4096, 20, synth.LEVELS4, dh.PROTO_NXDN, lambda k: synth.nxdn_symbols(16, seed=k, lead_in=0)),
    ("dstar", 8192, 10, synth.LEVELS2, dh.PROTO_DSTAR, lambda k: synth.dstar_symbols(60, seed=k, lead_in=0)),
    ("pocsag", 32768, 40, synth.LEVELS2[::-1].copy(), dh.PROTO_POCSAG,
     lambda k: synth.pocsag_bits([(1000 + k, 3, "THROUGHPUT")], seed=k, lead_in=0)),
]
for name, C, sps, levels, pid, gen in CASES:
    if gen is None:
        x, _ = synth.dmr_channel_bank(C, L, seed=1, device="cuda:0")
    else:
        nsym = L // sps + 8
        pool = np.stack([np.resize(gen(k), nsym) for k in range(16)])
        x = synth.modulate_batch(pool[np.arange(C) % 16], L, sps=sps, levels=levels, amplitude=0.5, snr_db=15.0,
                                 seed=1, device="cuda:0")
    pipe = dh.Pipe(C, pid, max_chunk=L)
    res = {}
    for mode in (False, True):
        pipe.set_async(mode)
        for _ in range(3):
            pipe.process(x, n=L)
            pipe.discard()
        pipe.sync()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            pipe.process(x, n=L)
            pipe.discard()
        pipe.sync()
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / steps
    pipe.set_async(False)
    print("%-7s %6d ch x %d: %.3f ms/step = %.1f Gsamples/s; pipelined %.3f ms/step = %.1f Gsamples/s" % (
        name, C, L, res[False], C * L / res[False] / 1e6, res[True], C * L / res[True] / 1e6))
    pipe.close()
    del x
