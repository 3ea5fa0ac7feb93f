"""Time one 48000-sample step of the DMR pipe when it is fed in smaller sequential chunks (L2-residency experiment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C, L = 4096, 48000
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
for chunk in (48000, 17408, 8704, 6528, 4352, 2176):
    pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
    def step():
        for pos in range(0, L, chunk):
            c = min(chunk, L - pos)
            pipe.process(x[:, pos:], n=c)
        pipe.decoder.discard()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("chunk %6d: %.3f ms/step  %.1f Gsamples/s" % (chunk, ms, C * L / ms / 1e6))
    pipe.close()
