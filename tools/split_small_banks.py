"""Step latency of the DMR pipe against the bank size for the two K2 schedules (dh_demod_set_split).
The one-kernel demodulator walks the 100-symbol blocks of a channel in order, so its duration does not shrink with
the bank; the split schedule only keeps the variance search on that chain.
usage: split_small_banks.py [steps] [channels,channels,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = 48000
banks = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 32, 256, 1024, 2048, 4096]
for C in banks:
    x, _ = synth.dmr_channel_bank(C, L, seed=1, device="cuda:0")
    row = []
    for split in (False, True):
        for mode in (False, True):
            pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
            pipe.set_demod_split(split)
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
            row.append(e0.elapsed_time(e1) / steps)
            pipe.close()
    print("%5d ch x %d: one kernel %.3f ms/step (pipelined %.3f) | split %.3f ms/step (pipelined %.3f)" % (
        C, L, row[0], row[1], row[2], row[3]))
