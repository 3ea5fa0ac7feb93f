"""Asynchronous (cross-call pipelined) steps inside one profiler range, for `ncu --replay-mode app-range`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C, L = 4096, 48000
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
use_async = (sys.argv[2] != "sync") if len(sys.argv) > 2 else True
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
pipe.set_async(use_async)
for _ in range(3):
    pipe.process(x, n=L)
    pipe.discard()
pipe.sync()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    pipe.process(x, n=L)
    pipe.discard()
pipe.sync()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
pipe.set_async(False)
print("done")
