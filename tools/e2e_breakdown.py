import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth
C, L = 4096, 48000
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
xh = torch.empty((C, pipe.host_pitch), dtype=torch.float32).pin_memory()
xh.copy_(x)
for _ in range(3):
    pipe.process(xh, n=L); pipe.collect(); pipe.decoder.clear()
tp = tc = tcl = 0
N = 10
for _ in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pipe.process(xh, n=L); torch.cuda.synchronize(); t1 = time.perf_counter()
    pipe.collect(); t2 = time.perf_counter()
    pipe.decoder.clear(); t3 = time.perf_counter()
    tp += t1 - t0; tc += t2 - t1; tcl += t3 - t2
print("process(H2D+kernels) %.2f ms, collect %.2f ms, clear %.2f ms" % (tp / N * 1e3, tc / N * 1e3, tcl / N * 1e3))
print("H2D GB/s if kernels 1.55ms: %.1f" % (C * pipe.host_pitch * 4 / ((tp / N) - 1.55e-3) / 1e9))
ev, d2h = pipe.decoder.stats(); print("events", ev, "d2h bytes", d2h)
