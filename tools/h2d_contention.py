"""Does the pinned host -> device copy slow down while the pipe's kernels run on the same GPU?  (Explains the gap between
e2e and the bare copy: tools/e2e_gap.py.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C, L = 4096, 48000
x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
pipe = dh.Pipe(C, dh.PROTO_DMR, max_chunk=L)
blk = dh.PinnedBlock(C, pipe.host_pitch_s16, dtype=torch.int16)
blk.tensor.zero_()
dst = torch.empty((C, pipe.host_pitch_s16), dtype=torch.int16, device="cuda:0")
copy_stream = torch.cuda.Stream()
work_stream = torch.cuda.Stream()


def copies(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        e0.record()
        for _ in range(k):
            dst.copy_(blk.tensor, non_blocking=True)
        e1.record()
    return e0, e1


for _ in range(3):
    pipe.process(x, n=L)
    pipe.discard()
torch.cuda.synchronize()
e0, e1 = copies(10)
torch.cuda.synchronize()
print("bare copy, idle GPU:              %.3f ms per 393 MB block" % (e0.elapsed_time(e1) / 10))
for duty, steps_per_copy in (("kernels back to back (100 %% busy)", 8), ("one step per copy (the e2e schedule, ~20 %% busy)", 1)):
    torch.cuda.synchronize()
    e0, e1 = copies(10)
    with torch.cuda.stream(work_stream):
        for _ in range(10 * steps_per_copy):
            pipe.process(x, n=L, stream=work_stream)
            pipe.discard(stream=work_stream)
            if steps_per_copy == 1:
                torch.cuda._sleep(int(5.5e-3 * 1.9e9))   # idle gap so that one step falls into each copy
    torch.cuda.synchronize()
    print("copy beside %-48s %.3f ms per block" % (duty + ":", e0.elapsed_time(e1) / 10))
# and the other direction at the same time (the result read-back of the e2e loop is a few MB per step)
back = torch.empty((C, 1024), dtype=torch.uint8).pin_memory()
src = torch.zeros((C, 1120), dtype=torch.uint8, device="cuda:0")
torch.cuda.synchronize()
e0, e1 = copies(10)
with torch.cuda.stream(work_stream):
    for _ in range(200):
        back.copy_(src[:, :1024], non_blocking=True)      # a 2-D device -> host copy with narrow rows
torch.cuda.synchronize()
print("copy beside 2-D device -> host copies (4096 rows x 1 KB, 200 of them): %.3f ms per block" % (e0.elapsed_time(e1) / 10))
