"""Short single-GPU run of the bench workload for ncu captures (never a bench value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
torch.cuda.synchronize()
for _ in range(steps):
    pipe.process(x, n=L)
    pipe.decoder.discard()
torch.cuda.synchronize()
print("done", pipe.launch_count)
