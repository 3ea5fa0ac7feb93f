"""N > 1 path on CPU (no GPU, no compute calls into the product): the host-side logic of the sharded pipe.

  * the channel partition the library computes (dh_shard_channel_range) tiles [0, N) for every world size;
  * the wire-block layout (dh_shard_wire_layout) is consistent across ranks and big enough for what a step can emit;
  * two gloo ranks model one scatter -> decode -> gather round trip on that layout: rank 0 scatters the sample rows
    of the library's ranges, each rank decodes its rows (with the CPU oracle — the GPU banks need a device) and fills
    a wire block exactly as shard.cu's pack kernel lays it out, rank 0 gathers the blocks and unpacks them in global
    channel order; the result must equal a single-process run.  The NCCL version of the same round trip, through
    the product, is tests/test_shard_gpu.py.
"""
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
    for channels in (1, 7, 8, 4096, 65536, 10, 65537):
        for world in (1, 2, 3, 8):
            if channels < world:
                continue
            spans = [shard.channel_range(r, world, channels) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == channels
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == sorted(sizes, reverse=True)      # the first ranks own the extra channels


def test_wire_layout_bounds():
    import digiham_b200 as dh
    for proto, sps, frame_syms, frame_bytes in ((dh.PROTO_DMR, 10, 144, 27), (dh.PROTO_YSF, 10, 480, 95),
                                                (dh.PROTO_NXDN, 20, 192, 36), (dh.PROTO_DSTAR, 10, 96, 9)):
        for n in (4800, 48000):
            w_out, w_ev, total = shard.wire_layout(proto, n, 1000)
            assert w_out % 16 == 0 and w_ev >= 1
            # what a step can emit at most: one frame's payload per frame_syms symbols of (chunk + carried tail)
            assert w_out >= frame_bytes * (n // (sps - 1) // frame_syms + 1)
            header = (3 * 1000 * 4 + 15) & ~15
            assert total == header + 1000 * w_out + 1000 * w_ev * 16
            # the block of a smaller shard is smaller and the layout is a pure function of its arguments
            assert shard.wire_layout(proto, n, 999)[2] < total
            assert shard.wire_layout(proto, n, 1000) == (w_out, w_ev, total)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(outs, w_out, w_ev):
    """numpy model of pack_results_kernel's block for byte streams only (events stay on the device path)."""
    n = len(outs)
    header = (3 * n * 4 + 15) & ~15
    blk = np.zeros(header + n * w_out + n * w_ev * 16, dtype=np.uint8)
    counts = blk[:3 * n * 4].view(np.uint32)
    for c, o in enumerate(outs):
        assert len(o) <= w_out
        counts[c] = len(o)
        blk[header + c * w_out: header + c * w_out + len(o)] = o
    return blk


def _unpack(blk, n, w_out):
    header = (3 * n * 4 + 15) & ~15
    counts = blk[:3 * n * 4].view(np.uint32)
    return [blk[header + c * w_out: header + c * w_out + int(counts[c])].tobytes() for c in range(n)]


def _worker(rank, world, port, channels, n, q):
    import digiham_b200 as dh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spans = [shard.channel_range(r, world, channels) for r in range(world)]
        lo, hi = spans[rank]
        pitch = (n + 3) & ~3
        x_full = None
        rows = None
        if rank == 0:
            x_full, _ = synth.dmr_channel_bank(channels, n, seed=21, device="cpu", noise_fraction=0.0)
            x_full = x_full[:, :pitch].contiguous()
            rows = [x_full[a:b].contiguous() for a, b in spans]
        x = torch.empty((hi - lo, pitch), dtype=torch.float32)
        if rank == 0:
            x.copy_(rows[0])
            for r in range(1, world):
                dist.send(rows[r], dst=r)
        else:
            dist.recv(x, src=0)
        orc = oracle_lib.best()
        _, outs, _ = orc.pipe_batch(oracle_lib.PROTO_DMR, x[:, :n].numpy(), threads=2)
        w_out, w_ev, total = shard.wire_layout(dh.PROTO_DMR, n, hi - lo)
        blk = torch.from_numpy(_pack(outs, w_out, w_ev))
        assert blk.numel() == total
        if rank == 0:
            blocks = [blk]
            for r in range(1, world):
                nb = shard.wire_layout(dh.PROTO_DMR, n, spans[r][1] - spans[r][0])[2]
                b = torch.empty(nb, dtype=torch.uint8)
                dist.recv(b, src=r)
                blocks.append(b)
            frames = []
            for r in range(world):
                frames.extend(_unpack(blocks[r].numpy(), spans[r][1] - spans[r][0], w_out))
            _, ref_outs, _ = orc.pipe_batch(oracle_lib.PROTO_DMR, x_full[:, :n].numpy(), threads=2)
            ok = len(frames) == channels and all(frames[c] == ref_outs[c].tobytes() for c in range(channels))
            q.put(("ok" if ok else "mismatch", sum(len(f) for f in frames)))
        else:
            dist.send(blk, dst=0)
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
