"""Pinned host -> device copy bandwidth of this box (context for bench.py's PCIe-bound e2e figure)."""
import torch
n = 4096 * 48000
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("H2D %d MB: %.3f ms = %.1f GB/s = %.2f Gsamples/s of float32" % (n * 4 // 1000000, ms, n * 4 / ms / 1e6, n / ms / 1e6))
