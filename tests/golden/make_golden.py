#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v1.npz (and golden_v2.npz: NXDN / D-Star; `make_golden.py v2` writes only that)
from the COMPILED REFERENCE (oracle/_ref, built from the unmodified
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

    # YSF decoder (all modes, symbol errors) and POCSAG decoder on bit streams; whole pipes on samples
    for k, (mode, err) in enumerate([("DN", 0.0), ("mix", 0.01), ("VW", 0.003)]):
        sym = synth.ysf_symbols(14, seed=200 + k, mode=mode, symbol_errors=err)
        o, m = ref.decode(oracle_lib.PROTO_YSF, sym)
        g["ysf%d_sym" % k] = sym
        g["ysf%d_out" % k] = o
        g["ysf%d_meta" % k] = np.frombuffer(m, dtype=np.uint8)
    bits = synth.pocsag_bits([(1234562, 3, "HELLO B200"), (77, 3, "THE QUICK BROWN FOX"), (9000, 1, "x"), (8, 3, "A")],
                             seed=1, bit_errors=2, lead_in=40)
    g["pocsag_bits"] = bits
    g["pocsag_out"] = ref.decode(oracle_lib.PROTO_POCSAG, bits)[0]
    xp = synth.modulate(bits[:1300], sps=40, levels=synth.LEVELS2[::-1].copy(), snr_db=15, ppm=200,
                        rng=np.random.default_rng(2))
    g["pocsag_pipe_in"] = xp
    ps, po, _ = ref.pipe(oracle_lib.PROTO_POCSAG, xp)
    g["pocsag_pipe_sym"] = ps
    g["pocsag_pipe_out"] = po
    ys = synth.ysf_symbols(6, seed=300, mode="DN", lead_in=30)
    xy = synth.modulate(ys, sps=10, snr_db=16, phase=5, rng=np.random.default_rng(3))
    g["ysf_pipe_in"] = xy
    _, yo, ym = ref.pipe(oracle_lib.PROTO_YSF, xy)
    g["ysf_pipe_out"] = yo
    g["ysf_pipe_meta"] = np.frombuffer(ym, dtype=np.uint8)

    # K6 and the YSF primitives
    a = rng.integers(-32768, 32768, 1500).astype(np.int16)
    a[:200] = (15000 * np.sin(np.arange(200) * 0.3)).astype(np.int16)
    g["dvf_in"] = a
    g["dvf_out"] = ref.dvf(a)
    for steps in (100, 180):
        packed = rng.integers(0, 256, size=(6, (steps + 3) // 4)).astype(np.uint8)
        # first rows: valid code sequences with a few dibit errors
        for r in range(3):
            enc = synth.ysf_conv_encode(rng.integers(0, 2, steps)).copy()
            for e in rng.choice(steps, size=r * 2, replace=False):
                enc[e] ^= int(rng.integers(1, 4))
            pk = np.zeros((steps + 3) // 4, dtype=np.uint8)
            for i in range(steps):
                pk[i // 4] |= enc[i] << (6 - 2 * (i % 4))
            packed[r] = pk
        res = []
        for r in range(6):
            metric, bits_out = ref.trellis(packed[r], steps)
            res.append(np.concatenate([[metric], bits_out]))
        g["trellis%d_in" % steps] = packed
        g["trellis%d_out" % steps] = np.stack(res).astype(np.uint8)
    blob = rng.integers(0, 256, 40).astype(np.uint8)
    g["crc_in"] = blob
    g["crc_out"] = np.array([ref.crc16(blob[:k]) for k in (4, 10, 20, 40)], dtype=np.uint32)
    g["whitening_out"] = ref.whitening(blob[:20], 160)

    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes,", len(g), "arrays")


def main_v2():
    """golden_v2.npz: NXDN and D-Star, same recipe as v1."""
    ref = oracle_lib.ref()
    assert ref is not None and ref.kind == "reference", "needs the compiled reference (make -C oracle ref)"
    g = {}
    rng = np.random.default_rng(20261018)
    for k, err in enumerate([0.0, 0.01, 0.04]):
        sym = synth.nxdn_symbols(60, seed=(404, 416, 420)[k], symbol_errors=err)
        o, m = ref.decode(oracle_lib.PROTO_NXDN, sym)
        g["nxdn%d_sym" % k] = sym
        g["nxdn%d_out" % k] = o
        g["nxdn%d_meta" % k] = np.frombuffer(m, dtype=np.uint8)
    xs = synth.modulate(synth.nxdn_symbols(14, seed=410, lead_in=40), sps=20, snr_db=16, phase=7,
                        rng=np.random.default_rng(5))
    g["nxdn_pipe_in"] = xs
    ps, po, pm = ref.pipe(oracle_lib.PROTO_NXDN, xs)
    g["nxdn_pipe_sym"] = ps
    g["nxdn_pipe_out"] = po
    g["nxdn_pipe_meta"] = np.frombuffer(pm, dtype=np.uint8)
    # Nxdn::Trellis on valid (punctured and clean) code sequences and on noise; SACCH / FACCH1 probes
    for nbits in (72, 192):
        packed = rng.integers(0, 256, size=(8, nbits // 8)).astype(np.uint8)
        for r in range(5):
            bits = np.concatenate([rng.integers(0, 2, nbits // 2 - 4), np.zeros(4, dtype=np.int64)]).astype(np.uint8)
            enc = synth.ysf_conv_encode(bits)
            coded = np.empty(nbits, dtype=np.uint8)
            coded[0::2] = enc >> 1
            coded[1::2] = enc & 1
            if r >= 2:
                coded[rng.choice(nbits, size=r, replace=False)] ^= 1
            packed[r] = np.packbits(coded)
        res = []
        for r in range(8):
            metric, out = ref.nxdn_trellis(packed[r], nbits)
            res.append(np.concatenate([[metric], out]))
        g["nxdn_trellis%d_in" % nbits] = packed
        g["nxdn_trellis%d_out" % nbits] = np.stack(res).astype(np.uint8)
    sac = np.stack([synth.nxdn_sacch_dibits(k % 4, 5 * k, rng.integers(0, 2, 18)) for k in range(12)])
    sac[8:] ^= (rng.random(sac[8:].shape) < 0.1).astype(np.uint8) * 2
    g["nxdn_sacch_in"] = sac
    g["nxdn_sacch_out"] = np.stack([np.concatenate([[int(ref.nxdn_sacch(d)[0])], ref.nxdn_sacch(d)[1] * int(ref.nxdn_sacch(d)[0])])
                                    for d in sac]).astype(np.uint8)
    fac = []
    for k in range(16):
        d = (rng.random(80) < (0.1 if k % 2 else 0.5)).astype(np.uint8)
        d[2:8] = synth._int_to_bits([0x08, 0x10, 0x01][k % 3], 6)
        fac.append(synth.nxdn_facch1_dibits(d))
    fac = np.stack(fac)
    g["nxdn_facch1_in"] = fac
    g["nxdn_facch1_out"] = np.array([ref.nxdn_facch1(d) for d in fac], dtype=np.int32)

    for k, err in enumerate([0.0, 0.004, 0.02]):
        sym = synth.dstar_symbols(260, seed=500 + k, bit_errors=err)
        o, m = ref.decode(oracle_lib.PROTO_DSTAR, sym)
        g["dstar%d_sym" % k] = sym
        g["dstar%d_out" % k] = o
        g["dstar%d_meta" % k] = np.frombuffer(m, dtype=np.uint8)
    ds = np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), 150), synth.dstar_symbols(50, seed=510, lead_in=0)])
    xd = synth.modulate(ds, sps=10, levels=synth.LEVELS2, snr_db=16, phase=3, rng=np.random.default_rng(6))
    g["dstar_pipe_in"] = xd
    ps, po, pm = ref.pipe(oracle_lib.PROTO_DSTAR, xd)
    g["dstar_pipe_sym"] = ps
    g["dstar_pipe_out"] = po
    g["dstar_pipe_meta"] = np.frombuffer(pm, dtype=np.uint8)
    hdrs, hres = [], []
    for k in range(10):
        hb = synth.dstar_header_bits(synth.dstar_header_bytes(flags=(0x80 if k == 3 else 0, 0, 0), my="GOLD%d" % k,
                                                              suffix=["", "B200"][k % 2])).copy()
        hb[rng.choice(660, size=[0, 2, 5, 9, 14, 20, 30, 45, 3, 7][k], replace=False)] ^= 1
        rc, text = ref.dstar_header(hb)
        hdrs.append(hb)
        hres.append(np.frombuffer(("%d|" % rc).encode() + text + b"\0" * (128 - len(text) - len("%d|" % rc)), dtype=np.uint8))
    g["dstar_header_in"] = np.stack(hdrs)
    g["dstar_header_out"] = np.stack(hres)

    path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "v2":
        main_v2()
    else:
        main()
        main_v2()
