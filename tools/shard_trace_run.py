"""Per-step timeline of the sharded pipe (DH_SHARD_TRACE=1): torchrun --nproc-per-node N tools/shard_trace_run.py [s16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import digiham_b200 as dh
from digiham_b200 import shard
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
s16 = len(sys.argv) > 1 and sys.argv[1] == "s16"
C, L = 8192, 48000
sp = shard.ShardedPipe(C * world, dh.PROTO_DMR, max_chunk=L, device=dev, fmt=dh.FMT_S16 if s16 else dh.FMT_F32)
blocks = None
if rank == 0:
    dt = torch.int16 if s16 else torch.float32
    blocks = [((torch.rand((C * world, sp.pitch), device=dev) - 0.5) * (20000 if s16 else 1)).to(dt) for _ in range(2)]
for i in range(10):
    sp.submit(blocks[i & 1] if rank == 0 else None, L, scatter=True)
    sp.discard_step()
sp.sync()
torch.cuda.synchronize()
sp.close()
dist.destroy_process_group()
