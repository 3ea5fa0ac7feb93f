"""Channel sharding across ranks (harness glue on top of torch.distributed).

The path shards trivially: channels are independent, rank r of R owns the contiguous range
[r*N/R, (r+1)*N/R) and no collective sits on the data path.  The two optional collectives move
data in and out when a single ingest rank holds all channels: `scatter_channels` (input sample blocks, NCCL
scatter over NVLink on GPUs, gloo on CPU) and `gather_frames` (decoded frames + metadata as fixed-slot byte rows).
"""
import torch
import torch.distributed as dist


def channel_range(rank, world, channels):
    """Contiguous, balanced split: the first (channels % world) ranks get one extra channel."""
    base, extra = divmod(channels, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def scatter_channels(x_full, channels, pitch, dtype=torch.float32, src=0, device="cpu"):
    """x_full: [channels, pitch] tensor on the ingest rank (None elsewhere).  Returns this rank's rows.
    Ranges may be ragged (channels % world != 0): rows are padded to the largest shard for the collective."""
    rank, world = _world()
    lo, hi = channel_range(rank, world, channels)
    if world == 1:
        return x_full[lo:hi]
    rows = max(channel_range(r, world, channels)[1] - channel_range(r, world, channels)[0] for r in range(world))
    recv = torch.empty((rows, pitch), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = channel_range(r, world, channels)
            c = x_full[a:b]
            if b - a < rows:
                pad = torch.zeros((rows - (b - a), pitch), dtype=dtype, device=device)
                c = torch.cat([c, pad], dim=0)
            chunks.append(c.contiguous())
    dist.scatter(recv, chunks, src=src)
    return recv[:hi - lo]


def pack_rows(items, width=None, device="cpu"):
    """list of bytes objects -> (uint8 tensor [n, width], int32 lengths); vectorised (no per-row Python work)."""
    import numpy as np
    lens_np = np.fromiter((len(b) for b in items), dtype=np.int64, count=len(items))
    if width is None:
        width = int(lens_np.max()) if len(items) else 0
    buf = np.zeros((len(items), max(1, width)), dtype=np.uint8)
    total = int(lens_np.sum())
    if total:
        flat = np.frombuffer(b"".join(bytes(b) for b in items), dtype=np.uint8)
        starts = np.cumsum(lens_np) - lens_np
        rows = np.repeat(np.arange(len(items)), lens_np)
        cols = np.arange(total) - np.repeat(starts, lens_np)
        buf[rows, cols] = flat
    return torch.from_numpy(buf).to(device), torch.from_numpy(lens_np.astype(np.int32)).to(device)


def gather_frames(local_items, channels, dst=0, device="cpu"):
    """local_items: list of bytes (one per local channel).  On `dst` returns the list for all channels in global
    channel order, elsewhere None.  Uses two collectives: all_reduce(MAX) for the slot width, gather for the rows."""
    rank, world = _world()
    if world == 1:
        return list(local_items)
    rows = max(channel_range(r, world, channels)[1] - channel_range(r, world, channels)[0] for r in range(world))
    width = torch.tensor([max([len(b) for b in local_items] + [1])], dtype=torch.int64, device=device)
    dist.all_reduce(width, op=dist.ReduceOp.MAX)
    width = int(width.item())
    padded = list(local_items) + [b""] * (rows - len(local_items))
    buf, lens = pack_rows(padded, width, device)
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    lenss = [torch.empty_like(lens) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst)
    dist.gather(lens, lenss, dst=dst)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = channel_range(r, world, channels)
        rb, rl = bufs[r].cpu().numpy(), lenss[r].cpu().numpy()
        raw = rb.tobytes()
        w = rb.shape[1]
        out.extend(raw[i * w:i * w + int(rl[i])] for i in range(b - a))
    return out
