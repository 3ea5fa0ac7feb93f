"""Quick K1 timing on the GPU box (not the bench contract; used while tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
for kind, name, nt in ((dh.RRC_WIDE, "wide", 81), (dh.RRC_NARROW, "narrow", 161)):
    x = torch.rand((C, n), device="cuda") - 0.5
    y = torch.empty_like(x)
    bank = dh.RrcBank(C, kind)
    for _ in range(3):
        bank.process(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        bank.process(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gs = C * n / ms / 1e6
    print("%s: %.3f ms  %.1f Gsamples/s  %.1f GB/s (8 B/sample)  %.2f TFLOP/s fp32 (%d flop/sample)" % (
        name, ms, gs, gs * 8, gs * 2 * nt / 1e3, 2 * nt))
