"""Randomised differential test of the decoder banks on SYMBOL streams: spliced valid traffic, truncated frames,
noise, bare sync words at random places, random chunking.  usage: fuzz_decoders.py [seconds] [seed] [max_rounds]   (max_rounds makes the run deterministic)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import digiham_b200 as dh
from digiham_b200 import synth
import oracle_lib

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
max_rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
rounds = 0
orc = oracle_lib.best()
rng = np.random.default_rng(seed0)

DMR_SYNCS = list(synth.DMR_SYNC.values())
PROTOS = {
    "dmr": (dh.PROTO_DMR, oracle_lib.PROTO_DMR, 4,
            lambda k: synth.dmr_symbols(12, seed=k, kinds=[("voice", "mixed"), ("mixed", "data")][k % 2], symbol_errors=[0, 0.01, 0.05][k % 3]),
            DMR_SYNCS),
    "ysf": (dh.PROTO_YSF, oracle_lib.PROTO_YSF, 4,
            lambda k: synth.ysf_symbols(5, seed=k, mode=["DN", "V1", "VW", "mix", "FR"][k % 5], symbol_errors=[0, 0.01, 0.05][k % 3]),
            [synth.YSF_SYNC]),
    "nxdn": (dh.PROTO_NXDN, oracle_lib.PROTO_NXDN, 4,
             lambda k: synth.nxdn_symbols(10, seed=k, symbol_errors=[0, 0.01, 0.05][k % 3]), [synth.NXDN_FSW]),
    "dstar": (dh.PROTO_DSTAR, oracle_lib.PROTO_DSTAR, 2,
              lambda k: synth.dstar_symbols(40, seed=k, bit_errors=[0, 0.003, 0.02][k % 3], gga=(k % 3 == 0)),
              [synth.DSTAR_HEADER_SYNC, synth.DSTAR_VOICE_SYNC, synth.DSTAR_TERMINATOR, synth.DSTAR_TERMINATOR[24:]]),
    "pocsag": (dh.PROTO_POCSAG, oracle_lib.PROTO_POCSAG, 2,
               lambda k: synth.pocsag_bits([(9 + k, 3, "DEC FUZZ %d" % k)], seed=k, bit_errors=k % 4, preamble=64 + (k % 7) * 32),
               [np.array([(synth.POCSAG_FSC >> (31 - i)) & 1 for i in range(32)], dtype=np.uint8)]),
}


def spliced(gen, levels, syncs, target):
    parts, total = [], 0
    while total < target:
        r = rng.random()
        if r < 0.45:
            s = gen(int(rng.integers(0, 1 << 30)))
            a = int(rng.integers(0, max(1, len(s) // 2)))
            b = int(rng.integers(a + 1, len(s) + 1))
            p = s[a:b]
        elif r < 0.75:
            p = rng.integers(0, levels, size=int(rng.integers(1, 600))).astype(np.uint8)
        else:
            p = syncs[int(rng.integers(0, len(syncs)))].copy()
            if rng.random() < 0.5:
                p[int(rng.integers(0, len(p)))] ^= 1
        parts.append(np.asarray(p, dtype=np.uint8))
        total += len(p)
    return np.concatenate(parts)[:target]


t_end = time.time() + budget
stats = {k: [0, 0] for k in PROTOS}
bad = 0
while time.time() < t_end and not bad:
    for name, (pid, oid, levels, gen, syncs) in PROTOS.items():
        C = int(rng.integers(1, 12))
        n = int(rng.integers(300, 9000))
        sym = np.stack([spliced(gen, levels, syncs, n) for _ in range(C)])
        bank = dh.DecoderBank(C, pid)
        if name == "dmr" and rng.random() < 0.5:
            filt = int(rng.integers(0, 4))
            bank.set_slot_filter(filt)
        else:
            filt = 3
        d = torch.from_numpy(sym).cuda()
        pos = 0
        while pos < n:
            c = int(min(n - pos, rng.integers(1, 2500)))
            # ragged: some channels receive fewer symbols in this call, the rest of their chunk follows in the next one
            bank.process(d[:, pos:pos + c].contiguous(), torch.full((C,), c, dtype=torch.int32, device="cuda"))
            bank.collect()
            pos += c
        for ch in range(C):
            ro, rm = orc.decode(oid, sym[ch], slot_filter=filt)
            if bank.output(ch) != ro.tobytes() or bank.meta(ch) != rm:
                print("MISMATCH proto=%s ch=%d C=%d n=%d filt=%d: %d vs %d bytes, meta %d vs %d" % (
                    name, ch, C, n, filt, len(bank.output(ch)), ro.size, len(bank.meta(ch)), len(rm)))
                np.save(os.path.join(ROOT, "gpurun_out", "fuzz_fail_%s.npy" % name), sym[ch])
                bad += 1
                break
            stats[name][1] += ro.size + len(rm)
        stats[name][0] += C
        bank.close()
        rounds += 1
        if bad or time.time() > t_end or rounds >= max_rounds:
            break
    if rounds >= max_rounds:
        break
for k, v in stats.items():
    print("%-7s channels %6d  bytes compared %9d" % (k, v[0], v[1]))
print("decoder fuzz: %s" % ("FAILED" if bad else "all equal"))
sys.exit(1 if bad else 0)
