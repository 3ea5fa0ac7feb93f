#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v1.npz from the COMPILED REFERENCE (oracle/_ref, built from the unmodified
sources under /root/reference by oracle/Makefile).  Run where the reference tree exists:

    python tests/golden/make_golden.py

The reference repository ships no tests or golden vectors of its own, so these fixtures — inputs
and the reference's outputs — are what pins the oracle on machines where /root/reference is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from digiham_b200 import synth  # noqa: E402


def main():
    ref = oracle_lib.ref()
    assert ref is not None and ref.kind == "reference", "needs the compiled reference (make -C oracle ref)"
    g = {}
    rng = np.random.default_rng(20261017)

    # K1: random samples incl. special values, wide + narrow
    x = rng.uniform(-1, 1, 700).astype(np.float32)
    x[5] = -0.0
    x[6] = 1e-40
    x[7] = 3e38
    g["rrc_in"] = x
    g["rrc_wide_out"] = ref.rrc(x, narrow=False)
    g["rrc_narrow_out"] = ref.rrc(x, narrow=True)

    # K2: RRC-filtered 4-level signal with a clock offset (timing loop steps), and raw 2-level at sps 40
    s4 = synth.random_symbols(900, 4, seed=3)
    y = ref.rrc(synth.modulate(s4, sps=10, ppm=700, phase=3, snr_db=14, rng=np.random.default_rng(4)))
    g["gfsk10_in"] = y
    g["gfsk10_out"] = ref.demod(y, sps=10, four_level=True)
    s2 = synth.random_symbols(400, 2, seed=5)
    z = synth.modulate(s2, sps=40, levels=synth.LEVELS2, ppm=-600, phase=11, snr_db=10, rng=np.random.default_rng(6))
    g["fsk40_in"] = z
    g["fsk40_inv_out"] = ref.demod(z, sps=40, four_level=False, invert=True)
    g["fsk40_out"] = ref.demod(z, sps=40, four_level=False, invert=False)

    # DMR decoder on symbols (with symbol errors) and the whole pipe on samples
    for k, (kinds, err) in enumerate([(("voice", "mixed"), 0.0), (("mixed", "data"), 0.02)]):
        sym = synth.dmr_symbols(70, seed=100 + k, kinds=kinds, symbol_errors=err)
        out, meta = ref.decode(oracle_lib.PROTO_DMR, sym)
        g["dmr%d_sym" % k] = sym
        g["dmr%d_out" % k] = out
        g["dmr%d_meta" % k] = np.frombuffer(meta, dtype=np.uint8)
    xb, _ = synth.dmr_channel_bank(2, 24000, seed=77, device="cpu", noise_fraction=0.0)
    for c in range(2):
        xc = xb[c, :24000].numpy()
        sym, out, meta = ref.pipe(oracle_lib.PROTO_DMR, xc)
        g["pipe%d_in" % c] = xc
        g["pipe%d_sym" % c] = sym
        g["pipe%d_out" % c] = out
        g["pipe%d_meta" % c] = np.frombuffer(meta, dtype=np.uint8)

    # block codes: decode result for every syndrome representative (low r bits) XOR a random codeword offset
    for cid, (name, r) in enumerate(zip(oracle_lib.FEC_NAMES, oracle_lib.FEC_PARITY_BITS)):
        res = np.zeros((1 << r, 2), dtype=np.uint32)
        for s in range(1 << r):
            ok, w = ref.fec(cid, s)
            res[s] = (int(ok), w)
        g["fec_" + name] = res
    pay = rng.integers(0, 256, size=(16, 25)).astype(np.uint8)
    for k in range(10):   # valid BPTC blocks with 0..4 bit errors; the rest stays random (mostly uncorrectable)
        bits = synth.dmr_bptc_encode(rng.integers(0, 256, size=12)).copy()
        for e in rng.choice(196, size=k % 5, replace=False):
            bits[e] ^= 1
        pay[k] = np.packbits(np.concatenate([bits, np.zeros(4, np.uint8)]))
    g["bptc_in"] = pay
    g["bptc_out"] = np.stack([np.concatenate([[int(ref.bptc(p)[0])], ref.bptc(p)[1]]) for p in pay]).astype(np.uint8)

    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
