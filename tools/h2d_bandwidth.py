"""Pinned host -> device copy bandwidth of this box (context for bench.py's PCIe-bound e2e figure):
regular pinned memory vs write-combined pinned memory."""
import ctypes
import torch

n = 4096 * 48000
nbytes = n * 4
d = torch.empty(n, dtype=torch.float32, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")


def timed(copy):
    for _ in range(3):
        copy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        copy()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


h = torch.empty(n, dtype=torch.float32).pin_memory()
ms = timed(lambda: d.copy_(h, non_blocking=True))
print("pinned         H2D %d MB: %.3f ms = %.1f GB/s = %.2f Gsamples/s" % (nbytes // 1000000, ms, nbytes / ms / 1e6, n / ms / 1e6))
for flag, name in ((4, "write-combined"), (0, "cudaHostAlloc default")):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flag))
    assert rc == 0, rc
    ctypes.memset(p, 0, nbytes)
    st = torch.cuda.current_stream().cuda_stream
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    ms = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), p, nbytes, 1, st))
    print("%-14s H2D %d MB: %.3f ms = %.1f GB/s = %.2f Gsamples/s" % (name, nbytes // 1000000, ms, nbytes / ms / 1e6, n / ms / 1e6))
    rt.cudaFreeHost(p)
