"""Tiny run of the three kernels of the split K2 schedule for compute-sanitizer (all sps variants, work rows and
caller rows, ragged chunks); not a test of values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C, L = 7, 9000
for sps, four in ((10, True), (20, True), (40, False), (12, False), (25, True)):
    L = 900 * sps
    sym = np.stack([synth.random_symbols(L // sps + 8, 4 if four else 2, k) for k in range(C)])
    x = synth.modulate_batch(sym, L, sps=sps, levels=synth.LEVELS4 if four else synth.LEVELS2, amplitude=0.5,
                             snr_db=12.0, seed=1, device="cuda:0")
    bank = dh.DemodBank(C, sps=sps, four_level=four)
    bank.set_split(True)
    pos = 0
    for c in (300 * sps, 1, 2 * sps + 1, 250 * sps + 3, L):
        c = min(c, L - pos)
        if c <= 0:
            break
        bank.process(x[:, pos:pos + c].contiguous())      # aligned caller rows are read in place, others are copied
        pos += c
    torch.cuda.synchronize()
    bank.close()
sym = np.stack([np.resize(synth.dmr_symbols(8, seed=k, lead_in=10), 908) for k in range(C)])
x = synth.modulate_batch(sym, 9000, sps=10, levels=synth.LEVELS4, amplitude=0.5, snr_db=12.0, seed=1, device="cuda:0")
for mode in (False, True):
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=4500)
    pipe.set_demod_split(True)
    pipe.set_async(mode)
    for pos in (0, 4500):
        pipe.process(x[:, pos:pos + 4500], n=4500)
        pipe.collect()
    pipe.set_async(False)
    pipe.close()
print("sanitize_split done")
