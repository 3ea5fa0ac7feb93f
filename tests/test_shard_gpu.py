"""Sharded pipe (dh_shard_*, BASELINE configs[3]): scatter -> kernels -> pack -> gather, parity-checked on the
gathering rank against the compiled reference.

  * world 1 on one GPU: the pack kernel, the wire layout and the root-side result sink, pipelined over several steps;
  * world 2 (skipped below two GPUs): two processes, NCCL scatter from rank 0 and gather to rank 0, float32 and int16
    ingest, ragged channel split; every channel of every rank is compared with the oracle.
"""
import os
import socket

import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_input(channels, n, steps, seed, s16, device):
    x, _ = synth.dmr_channel_bank(channels, n * steps, seed=seed, device=device)
    x = x[:, :n * steps]
    if s16:
        s = torch.clamp(torch.round(x * 20000.0), -32768, 32767).to(torch.int16)
        ref_in = s.cpu().numpy().astype(np.float32) / np.float32(32767)
        return s, ref_in
    return x, x.cpu().numpy()


def _run_sharded(sp, data, n, steps, pitch, scatter, rank_rows=None):
    """data: [channels, n * steps] on the root (or the local rows when not scattering)."""
    dev = sp.device
    blocks = []
    if data is not None:
        for k in range(steps):
            b = torch.zeros((data.shape[0], pitch), dtype=data.dtype, device=dev)
            b[:, :n] = data[:, k * n:(k + 1) * n]
            blocks.append(b)
    sp.submit(blocks[0] if blocks else None, n, scatter=scatter)
    for k in range(1, steps):
        sp.submit(blocks[k] if blocks else None, n, scatter=scatter)
        sp.collect_step()
    sp.collect_step()
    sp.sync()
    torch.cuda.synchronize()


@pytest.mark.parametrize("s16", [False, True])
def test_shard_world1_pack_and_sink(s16):
    import digiham_b200 as dh
    from digiham_b200 import shard
    C, n, steps = 70, 12000, 4
    data, ref_in = _make_input(C, n, steps, seed=41, s16=s16, device="cuda")
    sp = shard.ShardedPipe(C, dh.PROTO_DMR, max_chunk=n, device="cuda:0", fmt=dh.FMT_S16 if s16 else dh.FMT_F32)
    assert (sp.lo, sp.hi) == (0, C)
    _, wire_bytes, _ = sp.stats()
    assert wire_bytes == shard.wire_layout(dh.PROTO_DMR, n, C)[2]
    _run_sharded(sp, data, n, steps, sp.pitch, scatter=True)
    orc = oracle_lib.best()
    _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, ref_in, threads=8, chunk=4096)
    assert sum(len(o) for o in outs) > 27 * 50
    for c in range(C):
        assert sp.output(c) == outs[c].tobytes(), c
        assert sp.meta(c) == metas[c], c
    launches, _, d2h = sp.stats()
    # K1, K2, decoder, pack per step; K2 is three kernels when the bank is small enough for the split schedule
    assert launches in (steps * 4, steps * 6) and d2h > 0
    sp.close()


def _worker(rank, world, port, channels, n, steps, s16, q, no_ipc=False, root=0):
    if no_ipc:
        os.environ["DH_SHARD_NO_IPC"] = "1"
    else:
        os.environ.pop("DH_SHARD_NO_IPC", None)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import digiham_b200 as dh
    from digiham_b200 import shard
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sp = shard.ShardedPipe(channels, dh.PROTO_DMR, max_chunk=n, device=dev, fmt=dh.FMT_S16 if s16 else dh.FMT_F32,
                               root=root)
        data = ref_in = None
        if rank == root:
            data, ref_in = _make_input(channels, n, steps, seed=43, s16=s16, device=dev)
        _run_sharded(sp, data, n, steps, sp.pitch, scatter=True)
        if rank == root:
            orc = oracle_lib.best()
            _, outs, metas = orc.pipe_batch(oracle_lib.PROTO_DMR, ref_in, threads=8, chunk=4096)
            bad = [c for c in range(channels)
                   if sp.output(c) != outs[c].tobytes() or sp.meta(c) != metas[c]]
            q.put(("ok" if not bad else "mismatch %s" % bad[:8], sum(len(o) for o in outs), sp.hi - sp.lo,
                   sp.scatter_path * 10 + sp.gather_path))
        # second phase on the same object: every rank feeds its own rows (no scatter), results still gathered
        sp.clear()
        lo, hi = sp.lo, sp.hi
        local, local_ref = _make_input(hi - lo, n, 2, seed=100 + rank, s16=s16, device=dev)
        _run_sharded(sp, local, n, 2, sp.pitch, scatter=False)
        orc = oracle_lib.best()
        # the streams continue: the oracle sees phase-1 samples of these channels followed by the local ones
        gathered = [None] * world
        dist.all_gather_object(gathered, local_ref)
        if rank == root:
            full2 = np.concatenate(gathered, axis=0)
            _, outs2, metas2 = orc.pipe_batch(oracle_lib.PROTO_DMR, np.concatenate([ref_in, full2], axis=1), threads=8,
                                              chunk=4096)
            bad = []
            for c in range(channels):
                want_o = outs2[c].tobytes()[len(outs[c]):]
                want_m = metas2[c][len(metas[c]):]
                if sp.output(c) != want_o or sp.meta(c) != want_m:
                    bad.append(c)
            q.put(("ok" if not bad else "mismatch2 %s" % bad[:8], 0, 0))
        sp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("s16,no_ipc", [(False, False), (True, False), (False, True), (True, True)])
def test_shard_two_ranks_scatter_compute_gather(s16, no_ipc):
    """Both scatter paths: the root's copy engines writing into the IPC-mapped peer slots (default) and NCCL send / recv
    of the rows (DH_SHARD_NO_IPC=1, also the fallback where the slots cannot be mapped)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    channels, n, steps = 101, 12000, 4     # ragged split: 51 + 50 channels
    procs = [ctx.Process(target=_worker, args=(r, 2, port, channels, n, steps, s16, q, no_ipc)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    status, nbytes, nlocal, path = q.get(timeout=10)
    assert status == "ok" and nbytes > 27 * 50 and nlocal == 51
    assert path == (11 if no_ipc else 22), "scatter / gather paths %d" % path
    status2, _, _ = q.get(timeout=10)
    assert status2 == "ok"


def test_shard_two_ranks_root_is_not_rank_zero():
    """The ingest / gathering rank may be any rank (here rank 1 of 2)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 33, 12000, 3, True, q, False, 1)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    status, nbytes, nlocal, path = q.get(timeout=10)
    assert status == "ok" and nbytes > 0 and nlocal == 16 and path == 22
    assert q.get(timeout=10)[0] == "ok"


def _cpp_host(tmp_path, world):
    """examples/sharded_pipe.cpp: a C++ host (no Python, no torch in the processes) runs the sharded pipe through the
    C ABI alone; rank 0's gathered output must equal the reference for every channel."""
    import struct
    import subprocess
    root = oracle_lib.ROOT
    exe = str(tmp_path / "sharded_pipe")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include",
                    os.path.join(root, "examples", "sharded_pipe.cpp"), "-L" + os.path.join(root, "digiham_b200"),
                    "-ldigiham_b200", "-Wl,-rpath," + os.path.join(root, "digiham_b200"), "-L/usr/local/cuda/lib64",
                    "-lcudart", "-o", exe], check=True)
    channels, n, steps = 45, 12000, 3
    data, ref_in = _make_input(channels, n, steps, seed=47, s16=True, device="cuda")
    s = data.cpu().numpy()                                        # [channels, n * steps]
    blocks = np.stack([s[:, k * n:(k + 1) * n] for k in range(steps)])   # [steps][channels][n]
    inp = str(tmp_path / "in.s16")
    blocks.astype(np.int16).tofile(inp)
    prefix = str(tmp_path / "res")
    idf = str(tmp_path / "nccl.id")
    procs = [subprocess.Popen([exe, str(r), str(world), idf, inp, str(channels), str(n), str(steps), prefix],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    assert "%d channels over %d rank(s)" % (channels, world) in outs[0][0]
    _, ref_out, ref_meta = oracle_lib.best().pipe_batch(oracle_lib.PROTO_DMR, ref_in, threads=8, chunk=4096)

    def records(path):
        raw = open(path, "rb").read()
        pos, res = 0, []
        while pos < len(raw):
            (ln,) = struct.unpack_from("<I", raw, pos)
            res.append(raw[pos + 4:pos + 4 + ln])
            pos += 4 + ln
        return res

    got_out, got_meta = records(prefix + ".out"), records(prefix + ".meta")
    assert len(got_out) == channels and len(got_meta) == channels
    for c in range(channels):
        assert got_out[c] == ref_out[c].tobytes() and got_meta[c] == ref_meta[c], c
    assert sum(len(o) for o in got_out) > 27 * 20


def test_cpp_host_sharded_pipe_one_rank(tmp_path):
    _cpp_host(tmp_path, 1)


def test_cpp_host_sharded_pipe_two_ranks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _cpp_host(tmp_path, 2)


@pytest.mark.parametrize("name,s16", [("ysf", True), ("nxdn", False), ("dstar", False), ("pocsag", False)])
def test_shard_world1_other_protocols(name, s16):
    """The sharded pipe is protocol-agnostic (wire slots sized from each protocol's per-step bounds; the FSK pipes have no
    RRC stage, so their first kernel — the demodulator — is the reader of the scattered block)."""
    import bench
    import digiham_b200 as dh
    from digiham_b200 import shard
    C, n, steps = 40, 24000, 3
    dh_proto, orc_proto = bench.proto_ids(name)
    x = bench.workload_signal(name, C, n * steps, 5, "cuda")[:, :n * steps].contiguous()
    if s16:
        data = torch.clamp(torch.round(x * 20000.0), -32768, 32767).to(torch.int16)
        ref_in = data.cpu().numpy().astype(np.float32) / np.float32(32767)
    else:
        data, ref_in = x, x.cpu().numpy()
    sp = shard.ShardedPipe(C, dh_proto, max_chunk=n, device="cuda:0", fmt=dh.FMT_S16 if s16 else dh.FMT_F32)
    _run_sharded(sp, data, n, steps, sp.pitch, scatter=True)
    _, outs, metas = oracle_lib.best().pipe_batch(orc_proto, ref_in, threads=8, chunk=4096, meta_cap=1 << 15)
    assert sum(len(o) + len(m) for o, m in zip(outs, metas)) > 0
    for c in range(C):
        assert sp.output(c) == outs[c].tobytes() and sp.meta(c) == metas[c], (name, c)
    sp.close()
    if name == "pocsag":
        with pytest.raises(dh.DhError):      # no RRC stage to fuse the int16 conversion into
            bad = shard.ShardedPipe(C, dh_proto, max_chunk=n, device="cuda:0", fmt=dh.FMT_S16)
            blk = torch.zeros((C, bad.pitch), dtype=torch.int16, device="cuda")
            bad.submit(blk, n, scatter=True)


def test_shard_api_misuse_is_reported():
    """Error behaviour of dh_shard_*: return codes + dh_last_error, no crash, the object stays usable."""
    import digiham_b200 as dh
    from digiham_b200 import shard
    C, n = 8, 4000
    sp = shard.ShardedPipe(C, dh.PROTO_DMR, max_chunk=n, device="cuda:0")
    assert sp.scatter_path == 0
    blk = torch.zeros((C, sp.pitch), dtype=torch.float32, device="cuda")
    with pytest.raises(dh.DhError):
        sp.collect_step()                                   # nothing in flight
    with pytest.raises(dh.DhError):
        sp.submit(torch.zeros((C, sp.pitch + 4), dtype=torch.float32, device="cuda"), n)   # wrong pitch
    with pytest.raises(dh.DhError):
        sp.submit(blk, n + 1)                               # beyond max_chunk
    sp.submit(blk, n)
    sp.submit(blk, n)
    with pytest.raises(dh.DhError):
        sp.submit(blk, n)                                   # two steps already in flight
    sp.collect_step()
    sp.discard_step()
    with pytest.raises(dh.DhError):
        sp.output(C)                                        # channel out of range
    assert sp.output(0) == b"" and sp.meta(C - 1) == b""     # silence decodes to nothing
    sp.submit(blk, n)
    sp.collect_step()
    sp.close()
    with pytest.raises(dh.DhError):
        shard.ShardedPipe(0, dh.PROTO_DMR, max_chunk=n, device="cuda:0")     # fewer channels than ranks
    with pytest.raises(dh.DhError):
        shard.ShardedPipe(C, 99, max_chunk=n, device="cuda:0")               # unknown protocol
