"""Where the ~0.5 ms per step between e2e and the bare PCIe copy goes (streaming host interface, int16 blocks).
Variants: the bench loop as is; the same without clearing the host results; with 60 steps instead of 20 (fill / drain
amortised); host time of submit / collect_step per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C, L = 4096, 48000
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
for dtype in (torch.int16, torch.float32):
    s16 = dtype == torch.int16
    pitch = pipe.host_pitch_s16 if s16 else pipe.host_pitch
    src = torch.clamp(torch.round(x[:, :L] * 20000.0), -32768, 32767).to(torch.int16) if s16 else x[:, :L]
    blocks = [dh.PinnedBlock(C, pitch, dtype=dtype) for _ in range(2)]
    for b in blocks:
        b.tensor.zero_()
        b.tensor[:, :L].copy_(src)
    bufs = [b.tensor for b in blocks]
    link = torch.empty((C, pitch), dtype=dtype, device="cuda:0")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        link.copy_(bufs[0], non_blocking=True)
    torch.cuda.synchronize()
    link_ms = (time.perf_counter() - t0) / 10 * 1e3

    def loop(k, clear=True, timing=None):
        pipe.submit(bufs[0], n=L)
        for i in range(1, k):
            a = time.perf_counter()
            pipe.submit(bufs[i & 1], n=L)
            b = time.perf_counter()
            pipe.collect_step()
            c = time.perf_counter()
            if clear:
                pipe.decoder.clear()
            d = time.perf_counter()
            if timing is not None:
                timing.append((b - a, c - b, d - c))
        pipe.collect_step()
        pipe.decoder.clear()

    loop(3)
    for k, clear in ((20, True), (20, False), (60, True)):
        torch.cuda.synchronize()
        tm = []
        t0 = time.perf_counter()
        loop(k, clear, tm)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / k * 1e3
        sub = sum(t[0] for t in tm) / len(tm) * 1e3
        col = sum(t[1] for t in tm) / len(tm) * 1e3
        clr = sum(t[2] for t in tm) / len(tm) * 1e3
        print("%s  steps %2d clear=%-5s  %.3f ms/step (bare copy %.3f, ratio %.3f); host per step: submit %.3f, collect_step %.3f "
              "(blocks until the step is decoded), clear %.3f" % ("int16  " if s16 else "float32", k, clear, ms, link_ms,
                                                                 link_ms / ms, sub, col, clr))
    for b in blocks:
        b.close()
