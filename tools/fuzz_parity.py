"""Randomised differential test: every pipe against the CPU oracle with random sizes, chunkings and impairments.
usage: fuzz_parity.py [seconds] [seed] [max_rounds]   (GPU box; prints a summary line per protocol and exits 1 on a
mismatch; with max_rounds the run is deterministic: it stops after that many rounds, whatever the time budget)

D-Star streams with bit errors carry no NMEA GGA sentences: the REFERENCE aborts (std::stof throws std::invalid_argument,
src/dstar_decoder/dstar_phase.cpp:262-268) on a sentence whose fields were corrupted while its 8-bit checksum still
matches; the GPU path skips such a sentence.  Found by this tool (seed 101)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import digiham_b200 as dh
from digiham_b200 import synth
import oracle_lib

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
max_rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
orc = oracle_lib.best()
PROTOS = {
    "dmr": (dh.PROTO_DMR, oracle_lib.PROTO_DMR, 10, synth.LEVELS4,
            lambda k, e: synth.dmr_symbols(30, seed=k, kinds=[("voice", "mixed"), ("mixed", "data"), ("idle", "voice")][k % 3],
                                           symbol_errors=e)),
    "ysf": (dh.PROTO_YSF, oracle_lib.PROTO_YSF, 10, synth.LEVELS4,
            lambda k, e: synth.ysf_symbols(10, seed=k, mode=["DN", "V1", "VW", "mix", "FR"][k % 5], symbol_errors=e)),
    "nxdn": (dh.PROTO_NXDN, oracle_lib.PROTO_NXDN, 20, synth.LEVELS4,
             lambda k, e: synth.nxdn_symbols(25, seed=k, symbol_errors=e)),
    "dstar": (dh.PROTO_DSTAR, oracle_lib.PROTO_DSTAR, 10, synth.LEVELS2,
              lambda k, e: np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), 120),
                                           synth.dstar_symbols(80, seed=k, bit_errors=e, gga=(e == 0.0))])),
    "pocsag": (dh.PROTO_POCSAG, oracle_lib.PROTO_POCSAG, 40, synth.LEVELS2[::-1].copy(),
               lambda k, e: synth.pocsag_bits([(100 + k, 3, "FUZZ %d" % k), (7 + k, [0, 3][k % 2], "12345 6789")], seed=k,
                                              bit_errors=k % 4, lead_in=k % 50, preamble=int(100 + 50 * (k % 9)))),
}
rng = np.random.default_rng(seed0)
t_end = time.time() + budget
stats = {k: [0, 0, 0, 0, 0] for k in PROTOS}   # rounds, channels, bytes compared, migrations, int16 rounds
bad = 0
rnd = 0
while time.time() < t_end and not bad:
    for name, (pid, oid, sps, levels, gen) in PROTOS.items():
        rnd += 1
        C = int(rng.integers(1, 24))
        errs = rng.choice([0.0, 0.002, 0.01, 0.04], size=C)
        streams = [gen(int(rng.integers(0, 1 << 30)), float(errs[c])) for c in range(C)]
        nsym = min(len(s) for s in streams)
        nsym = int(rng.integers(max(60, nsym // 3), nsym + 1))
        sym = np.stack([s[:nsym] for s in streams])
        n = nsym * sps - int(rng.integers(0, sps))
        x = synth.modulate_batch(sym, n, sps=sps, levels=levels, amplitude=rng.choice([0.1, 0.5, 0.9], size=C),
                                 ppm=rng.choice([0.0, 40.0, -40.0, 90.0], size=C), phase=rng.integers(0, 3 * sps, size=C).astype(np.float64),
                                 snr_db=rng.choice([np.inf, 25.0, 12.0, 7.0], size=C), dc=rng.choice([0.0, 0.03], size=C),
                                 seed=int(rng.integers(0, 1 << 30)), device="cuda")
        max_chunk = int(rng.choice([1000, 4096, 12000, 48000]))
        pipe = dh.Pipe(C, pid, max_chunk=max_chunk)
        use_async = bool(rng.integers(0, 2))
        pipe.set_async(use_async)
        # r02: int16 ingest on the pipes with an RRC stage (csdr convert fused into K1), and a mid-stream migration
        # of all channels to a new pipe through a state blob
        use_s16 = name in ("dmr", "ysf", "nxdn") and bool(rng.integers(0, 2))
        if use_s16:
            x = torch.clamp(torch.round(x[:, :n] * 20000.0), -32768, 32767).to(torch.int16)
            ref_in = x.cpu().numpy().astype(np.float32) / np.float32(32767)
        else:
            ref_in = x[:, :n].cpu().numpy()
        migrate_at = int(rng.integers(1, 6)) if rng.random() < 0.35 else -1
        got = [[b"", b""] for _ in range(C)]

        def drain():
            pipe.collect()
            for ch in range(C):
                got[ch][0] += pipe.output(ch)
                got[ch][1] += pipe.meta(ch)
            pipe.decoder.clear()

        pos, calls = 0, 0
        unit = 8 if use_s16 else 4
        while pos < n:
            c = int(min(n - pos, rng.integers(1, max_chunk + 1)))
            blk = torch.zeros((C, (c + unit - 1) // unit * unit), dtype=x.dtype, device="cuda")
            blk[:, :c] = x[:, pos:pos + c]
            pipe.process(blk, n=c)
            pipe.sync()          # blk goes out of scope
            calls += 1
            if calls % 2 == 0:
                drain()
            pos += c
            if calls == migrate_at:
                drain()
                blob = pipe.export_state()
                pipe.close()
                pipe = dh.Pipe(C, pid, max_chunk=max_chunk)
                pipe.set_async(use_async)
                pipe.import_state(blob)
                stats[name][3] += 1
        drain()
        pipe.set_async(False)
        _, outs, metas = orc.pipe_batch(oid, ref_in, threads=8, meta_cap=1 << 16)
        for ch in range(C):
            if got[ch][0] != outs[ch].tobytes() or got[ch][1] != metas[ch]:
                print("MISMATCH proto=%s round=%d ch=%d C=%d n=%d max_chunk=%d async=%s s16=%s migrate=%d: %d vs %d bytes, meta %d vs %d" % (
                    name, rnd, ch, C, n, max_chunk, use_async, use_s16, migrate_at, len(got[ch][0]), outs[ch].size,
                    len(got[ch][1]), len(metas[ch])))
                bad += 1
                break
            stats[name][2] += outs[ch].size + len(metas[ch])
        stats[name][4] += int(use_s16)
        stats[name][0] += 1
        stats[name][1] += C
        pipe.close()
        if bad or time.time() > t_end or rnd >= max_rounds:
            break
    if rnd >= max_rounds:
        break
for k, v in stats.items():
    print("%-7s rounds %4d  channels %6d  bytes compared %9d  state migrations %3d  int16 rounds %3d" % (k, v[0], v[1], v[2], v[3], v[4]))
print("fuzz: %s" % ("FAILED" if bad else "all equal"))
sys.exit(1 if bad else 0)
