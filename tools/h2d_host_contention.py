"""Bare pinned host -> device copies while host threads are busy (the e2e loop replays metadata on 8 threads per step)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import digiham_b200 as dh

C, P = 4096, 48000
src = dh.PinnedBlock(C, P, dtype=torch.int16)
src.tensor.zero_()
dst = torch.empty((C, P), dtype=torch.int16, device="cuda:0")


def copies(k=12):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        dst.copy_(src.tensor, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


copies(3)
print("idle host:                          %.3f ms per 393 MB block" % copies())
stop = False


def burn_mem():
    a = np.zeros(8 << 20, dtype=np.uint8)
    b = np.empty_like(a)
    while not stop:
        np.copyto(b, a)


def burn_alu():
    x = np.random.default_rng(0).random(4096)
    while not stop:
        x = np.sqrt(x * x + 1.0)


for name, fn, nthreads in (("8 threads streaming memory", burn_mem, 8), ("8 threads of arithmetic", burn_alu, 8),
                           ("2 threads streaming memory", burn_mem, 2)):
    stop = False
    ts = [threading.Thread(target=fn) for _ in range(nthreads)]
    for t in ts:
        t.start()
    time.sleep(0.2)
    ms = copies()
    stop = True
    for t in ts:
        t.join()
    print("%-35s %.3f ms per block" % (name + ":", ms))
