"""Channel sharding across ranks: ctypes glue over the dh_shard_* C ABI (include/digiham_b200.h).

The path shards trivially: channels are independent, rank r of R owns the contiguous range
[r*N/R, (r+1)*N/R).  The data-path collectives — input scatter from the ingest rank, gather of the
decoded frames + metadata events — live in the library (digiham_b200/csrc/shard.cu: grouped ncclSend / ncclRecv on
device buffers, pipelined with the kernels).  torch.distributed is used for ONE thing here: handing rank 0's
128-byte NCCL unique id to the other ranks.
"""
import ctypes

import torch
import torch.distributed as dist

from ._capi import FMT_F32, FMT_S16, SHARD_SCATTER, PROTO_DMR, Pipe, DecoderBank, check, lib, _dev_index, _stream_ptr


def channel_range(rank, world, channels):
    """[lo, hi) of `rank`: contiguous, balanced, the first (channels % world) ranks get one extra channel.
    Computed by the library (dh_shard_channel_range, host-only)."""
    lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
    check(lib().dh_shard_channel_range(int(channels), int(world), int(rank), ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def wire_layout(proto, max_chunk, channels):
    """(bytes per channel slot, event records per channel slot, bytes of the whole wire block); host-only."""
    a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t()
    check(lib().dh_shard_wire_layout(int(proto), int(max_chunk), int(channels), ctypes.byref(a), ctypes.byref(b),
                                     ctypes.byref(c)))
    return a.value, b.value, c.value


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def make_comm(device):
    """A fresh NCCL communicator over all ranks of the default process group, created by the library
    (dh_shard_unique_id on rank 0 -> broadcast -> dh_shard_comm_init).  Returns an opaque pointer (None at world 1)."""
    rank, world = _world()
    if world == 1:
        return None
    uid = (ctypes.c_uint8 * 128)()
    if rank == 0:
        check(lib().dh_shard_unique_id(uid))
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    uid = (ctypes.c_uint8 * 128).from_buffer_copy(box[0])
    comm = ctypes.c_void_p()
    check(lib().dh_shard_comm_init(ctypes.byref(comm), uid, rank, world, _dev_index(device)))
    return comm


def destroy_comm(comm):
    if comm:
        check(lib().dh_shard_comm_destroy(comm))


class ShardedPipe:
    """One protocol pipe over all ranks (dh_shard_*): every method is collective."""

    def __init__(self, channels_total, proto=PROTO_DMR, max_chunk=48000, device="cuda:0", fmt=FMT_F32, root=0,
                 comm=None):
        self.rank, self.world = _world()
        self.root = root
        self.channels_total = int(channels_total)
        self.device = torch.device(device)
        self.fmt = fmt
        self._own_comm = comm is None and self.world > 1
        self._comm = make_comm(device) if self._own_comm else comm
        self._h = ctypes.c_void_p()
        check(lib().dh_shard_create(ctypes.byref(self._h), self._comm, self.rank, self.world, root, _dev_index(device),
                                    self.channels_total, proto, int(max_chunk), fmt))
        self.lo, self.hi = channel_range(self.rank, self.world, self.channels_total)
        # a non-owning view of this rank's pipe
        self.pipe = Pipe.__new__(Pipe)
        self.pipe._h = ctypes.c_void_p(lib().dh_shard_pipe(self._h))
        self.pipe.channels = self.hi - self.lo
        self.pipe.proto = proto
        self.pipe.max_chunk = int(max_chunk)
        self.pipe.device = self.device
        self.pipe.close = lambda: None
        self.pipe.decoder = DecoderBank.__new__(DecoderBank)
        self.pipe.decoder._h = ctypes.c_void_p(lib().dh_pipe_decoder(self.pipe._h))
        self.pipe.decoder.channels = self.pipe.channels
        self.pipe.decoder.close = lambda: None

    @property
    def pitch(self):
        return lib().dh_shard_pitch(self._h)

    @property
    def is_root(self):
        return self.rank == self.root

    def submit(self, x, n, scatter=True, stream=None):
        """x: the device block ([channels_total, pitch] on the root when scattering, None elsewhere; this rank's
        [local channels, pitch] otherwise), float32 or int16 according to the shard's format."""
        ptr, pitch = (x.data_ptr(), x.stride(0)) if x is not None else (None, 0)
        if x is not None:
            assert x.is_cuda and x.stride(1) == 1
            assert x.dtype == (torch.int16 if self.fmt == FMT_S16 else torch.float32)
        check(lib().dh_shard_submit_device(self._h, ptr, pitch, int(n), SHARD_SCATTER if scatter else 0,
                                           _stream_ptr(stream)))

    def collect_step(self):
        check(lib().dh_shard_collect_step(self._h))

    def discard_step(self):
        check(lib().dh_shard_discard_step(self._h))

    def sync(self, stream=None):
        check(lib().dh_shard_sync(self._h, _stream_ptr(stream)))

    def output(self, channel):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        check(lib().dh_shard_output(self._h, int(channel), ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value) if n.value else b""

    def meta(self, channel):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        check(lib().dh_shard_meta(self._h, int(channel), ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value) if n.value else b""

    def clear(self):
        check(lib().dh_shard_clear(self._h))

    @property
    def scatter_path(self):
        """0 = single rank, 1 = NCCL send / recv, 2 = copy engines into IPC-mapped peer slots."""
        return lib().dh_shard_scatter_path(self._h)

    @property
    def gather_path(self):
        """0 = single rank, 1 = NCCL send / recv, 2 = pack kernel stores into the root's IPC-mapped wire buffer."""
        return lib().dh_shard_gather_path(self._h)

    def stats(self):
        """(kernels launched by this rank, bytes of its wire block per step, bytes read back on the root)."""
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().dh_shard_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def close(self):
        if self._h:
            lib().dh_shard_destroy(self._h)
            self._h = ctypes.c_void_p()
        if self._own_comm and self._comm:
            destroy_comm(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
