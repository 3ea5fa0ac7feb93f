import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
import digiham_b200 as dh
from digiham_b200 import synth
C, L = 4096, 48000
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
pitch = pipe.host_pitch_s16
src = torch.clamp(torch.round(x[:, :L] * 20000.0), -32768, 32767).to(torch.int16)
blocks = [dh.PinnedBlock(C, pitch, dtype=torch.int16) for _ in range(2)]
for b in blocks:
    b.tensor.zero_(); b.tensor[:, :L].copy_(src)
bufs = [b.tensor for b in blocks]
k = 12
pipe.submit(bufs[0], n=L)
for i in range(1, k):
    pipe.submit(bufs[i & 1], n=L)
    pipe.collect_step()
    pipe.decoder.clear()
pipe.collect_step()
pipe.close()
