"""Short single-GPU run of a pipe workload for ncu captures (never a bench value).
usage: ncu_driver.py [channels] [samples] [steps] [proto: dmr|ysf|nxdn|dstar|pocsag] [mode: f32|s16|shard]
mode s16: int16 blocks (csdr convert fused into K1); mode shard: a world-1 dh_shard (adds the wire pack kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import digiham_b200 as dh
from digiham_b200 import synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
proto = sys.argv[4] if len(sys.argv) > 4 else "dmr"
if proto == "dmr":
    x, _ = synth.dmr_channel_bank(C, L, seed=1234, device="cuda:0")
    pid = dh.PROTO_DMR
else:
    sps, levels, pid, gen = {
        "ysf": (10, synth.LEVELS4, dh.PROTO_YSF, lambda k: synth.ysf_symbols(12, seed=k, mode="mix", lead_in=0)),
        "nxdn": (20, synth.LEVELS4, dh.PROTO_NXDN, lambda k: synth.nxdn_symbols(16, seed=k, lead_in=0)),
        "dstar": (10, synth.LEVELS2, dh.PROTO_DSTAR, lambda k: synth.dstar_symbols(60, seed=k, lead_in=0)),
        "pocsag": (40, synth.LEVELS2[::-1].copy(), dh.PROTO_POCSAG,
                   lambda k: synth.pocsag_bits([(1000 + k, 3, "NCU CAPTURE")], seed=k, lead_in=0)),
    }[proto]
    nsym = L // sps + 8
    pool = np.stack([np.resize(gen(k), nsym) for k in range(16)])
    sym = pool[np.arange(C) % 16]
    x = synth.modulate_batch(sym, L, sps=sps, levels=levels, amplitude=0.5, snr_db=15.0, seed=1, device="cuda:0")
mode = sys.argv[5] if len(sys.argv) > 5 else "f32"
if mode in ("s16", "shard"):
    pitch = (L + 7) & ~7
    xs = torch.zeros((C, pitch), dtype=torch.int16, device="cuda:0")
    xs[:, :L] = torch.clamp(torch.round(x[:, :L] * 20000.0), -32768, 32767).to(torch.int16)
    x = xs
if mode == "shard":
    from digiham_b200 import shard
    sp = shard.ShardedPipe(C, pid, max_chunk=L, device="cuda:0", fmt=dh.FMT_S16)
    torch.cuda.synchronize()
    for _ in range(steps):
        sp.submit(x, L, scatter=True)
        sp.discard_step()
    sp.sync()
    torch.cuda.synchronize()
    print("done", sp.stats())
else:
    pipe = dh.Pipe(C, pid, max_chunk=L)
    torch.cuda.synchronize()
    for _ in range(steps):
        pipe.process(x, n=L)
        pipe.decoder.discard()
    torch.cuda.synchronize()
    print("done", pipe.launch_count)
