"""Bare pinned host -> device copies: one block repeatedly vs two alternating blocks (what the streaming interface does)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh

C, P = 4096, 48000
for dtype, name in ((torch.int16, "int16"), (torch.float32, "float32")):
    src = [dh.PinnedBlock(C, P, dtype=dtype) for _ in range(3)]
    for b in src:
        b.tensor.zero_()
    dst = [torch.empty((C, P), dtype=dtype, device="cuda:0") for _ in range(3)]
    def run(nsrc, ndst, k=12):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            dst[i % ndst].copy_(src[i % nsrc].tensor, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k
    run(1, 1)
    for nsrc, ndst in ((1, 1), (2, 1), (1, 2), (2, 2), (3, 3)):
        ms = run(nsrc, ndst)
        print("%-7s %d host block(s) -> %d device block(s): %.3f ms per block, %.2f GB/s" % (
            name, nsrc, ndst, ms, C * P * src[0].tensor.element_size() / ms / 1e6))
    for b in src:
        b.close()
