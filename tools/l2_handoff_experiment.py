"""K1 -> K2 hand-off through L2 instead of DRAM: an upper bound for what an on-chip (fused) hand-off can buy in
memory terms, measured with the existing kernels.

A fused K1+K2 kernel would keep the filtered tile in shared memory; the only thing that changes for the memory
system is that K1's 0.79 GB of output per step is never written to / re-read from DRAM.  The same effect is had
without writing the fused kernel by cutting the 4096 channels into sub-banks whose filtered block fits the 126 MB L2
(512 channels x 48000 samples x 4 B = 98 MB) and running K1, K2, K3 of one sub-bank back to back: K2's reads then hit L2.
usage: l2_handoff_experiment.py [channels] [samples] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
stream = torch.cuda.current_stream()


def run(sub, use_async):
    pipes = [dh.Pipe(sub, dh.PROTO_DMR, max_chunk=L) for _ in range(C // sub)]
    for p in pipes:
        p.set_async(use_async)

    def step():
        for k, p in enumerate(pipes):
            p.process(x[k * sub:(k + 1) * sub], n=L)
            p.discard()

    for _ in range(3):
        step()
    for p in pipes:
        p.sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    for p in pipes:
        p.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    for p in pipes:
        p.close()
    return ms


print("channels %d x %d samples, %d steps; filtered block of a sub-bank: sub x %d x 4 B" % (C, L, steps, L))
for use_async in (False, True):
    for sub in (C, C // 2, C // 4, C // 8, C // 16):
        ms = run(sub, use_async)
        print("%-22s sub-bank %5d ch (%6.1f MB filtered)  %.4f ms/step  %.1f Gsamples/s" % (
            "cross-step pipelined" if use_async else "back to back", sub, sub * L * 4 / 1e6, ms, C * L / ms / 1e6))
