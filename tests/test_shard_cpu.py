"""N > 1 path on CPU: two gloo ranks shard the channels, rank 0 scatters the sample blocks, every rank runs its
channel range (here through the CPU oracle — the GPU banks need a device), rank 0 gathers the decoded frames and
metadata and compares them with a single-process run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib
from digiham_b200 import shard, synth


def test_channel_range_partitions():
    for channels in (1, 7, 8, 4096, 65536, 10):
        for world in (1, 2, 3, 8):
            spans = [shard.channel_range(r, world, channels) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == channels
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, channels, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x_full = None
        if rank == 0:
            x_full, _ = synth.dmr_channel_bank(channels, n, seed=21, device="cpu", noise_fraction=0.0)
        pitch = (n + 3) & ~3
        x = shard.scatter_channels(x_full, channels, pitch, device="cpu")
        lo, hi = shard.channel_range(rank, world, channels)
        assert x.shape[0] == hi - lo
        orc = oracle_lib.best()
        _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, x[:, :n].numpy(), threads=2)
        frames = shard.gather_frames([o.tobytes() for o in outs], channels)
        meta = shard.gather_frames(metas, channels)
        if rank == 0:
            _, ref_outs, ref_metas = orc.pipe_batch(oracle_lib.PROTO_DMR, x_full[:, :n].numpy(), threads=2)
            ok = all(frames[c] == ref_outs[c].tobytes() and meta[c] == ref_metas[c] for c in range(channels))
            q.put(("ok" if ok else "mismatch", sum(len(f) for f in frames)))
    finally:
        dist.destroy_process_group()


def test_two_rank_scatter_process_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    channels, n = 7, 30000      # ragged split: 4 + 3 channels
    procs = [ctx.Process(target=_worker, args=(r, 2, port, channels, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    status, nbytes = q.get(timeout=10)
    assert status == "ok" and nbytes > 0
