"""K3/K4 parity for DMR: dh_decoder_* (CUDA through the C ABI) vs the CPU oracle's Dmr::Decoder.

Byte stream (27-byte voice frames) byte-exact, metadata lines string-exact in order.  Streams come from the
seeded DMR base-station generator (valid TACT / Golay / BPTC / EMB / embedded LC, talker alias, GPS) with random
symbol errors so that correction, failure and sync-loss branches all run.
"""
import numpy as np
import pytest
import torch

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

KINDS = [("voice", "mixed"), ("mixed", "voice"), ("idle", "voice"), ("data", "mixed"), ("voice", "voice"),
         ("mixed", "data"), ("idle", "idle")]


def _streams(C, frames, seed, err_levels=(0.0, 0.002, 0.01, 0.03, 0.08)):
    out = []
    for ch in range(C):
        s = synth.dmr_symbols(frames, seed=seed * 100 + ch, kinds=KINDS[ch % len(KINDS)],
                              symbol_errors=err_levels[ch % len(err_levels)])
        out.append(s)
    n = min(len(s) for s in out)
    return np.stack([s[:n] for s in out])


def _gpu_decode(bank, sym, chunks):
    C = sym.shape[0]
    pos = 0
    for c in chunks:
        blk = torch.from_numpy(np.ascontiguousarray(sym[:, pos:pos + c])).cuda()
        nsym = torch.full((C,), c, dtype=torch.int32, device="cuda")
        bank.process(blk, nsym)
        bank.collect()
        pos += c
    return [bank.output(ch) for ch in range(C)], [bank.meta(ch) for ch in range(C)]


def _check(got_out, got_meta, sym, slot_filter=3, chunk=0):
    orc = oracle_lib.best()
    for ch in range(sym.shape[0]):
        ref_out, ref_meta = orc.decode(oracle_lib.PROTO_DMR, sym[ch], chunk=chunk, slot_filter=slot_filter)
        assert got_out[ch] == ref_out.tobytes(), "channel %d: voice bytes differ (%d vs %d)" % (
            ch, len(got_out[ch]), ref_out.size)
        assert got_meta[ch] == ref_meta, "channel %d meta differs:\n%s\n---\n%s" % (
            ch, got_meta[ch].decode(errors="replace")[:600], ref_meta.decode(errors="replace")[:600])


def test_dmr_decoder_whole_stream():
    import digiham_b200 as dh
    C = 35
    sym = _streams(C, 260, seed=1)
    bank = dh.DecoderBank(C, dh.PROTO_DMR)
    out, meta = _gpu_decode(bank, sym, [sym.shape[1]])
    _check(out, meta, sym)
    assert sum(len(o) for o in out) > 27 * 100 and sum(len(m) for m in meta) > 1000
    bank.close()


def test_dmr_decoder_streaming_chunks():
    import digiham_b200 as dh
    C = 10
    sym = _streams(C, 150, seed=2)
    n = sym.shape[1]
    rng = np.random.default_rng(9)
    chunks, left = [], n
    while left > 0:
        c = int(min(left, rng.choice([1, 13, 90, 91, 143, 144, 145, 480, 1000, 4800])))
        chunks.append(c)
        left -= c
    bank = dh.DecoderBank(C, dh.PROTO_DMR)
    out, meta = _gpu_decode(bank, sym, chunks)
    _check(out, meta, sym, chunk=128)
    bank.close()


@pytest.mark.parametrize("slot_filter", [0, 1, 2])
def test_dmr_decoder_slot_filter(slot_filter):
    import digiham_b200 as dh
    C = 7
    sym = _streams(C, 120, seed=3, err_levels=(0.0, 0.01))
    bank = dh.DecoderBank(C, dh.PROTO_DMR)
    bank.set_slot_filter(slot_filter)
    out, meta = _gpu_decode(bank, sym, [sym.shape[1]])
    _check(out, meta, sym, slot_filter=slot_filter)
    bank.close()


def test_dmr_decoder_noise_and_ragged_counts():
    """Pure noise never syncs; channels may receive different symbol counts per call."""
    import digiham_b200 as dh
    rng = np.random.default_rng(4)
    C = 6
    n = 9000
    sym = rng.integers(0, 4, size=(C, n)).astype(np.uint8)
    good = _streams(3, 60, seed=5, err_levels=(0.0,))
    sym[:3, :good.shape[1]] = good[:, :n]
    lens = [n, n - 1, n - 77, 100, 0, 5000]
    bank = dh.DecoderBank(C, dh.PROTO_DMR)
    blk = torch.from_numpy(sym).cuda()
    nsym = torch.tensor(lens, dtype=torch.int32, device="cuda")
    bank.process(blk, nsym)
    bank.collect()
    orc = oracle_lib.best()
    for ch in range(C):
        ref_out, ref_meta = orc.decode(oracle_lib.PROTO_DMR, sym[ch, :lens[ch]])
        assert bank.output(ch) == ref_out.tobytes(), ch
        assert bank.meta(ch) == ref_meta, ch
    bank.close()


# ---- opt-in mode beyond the reference: Reed-Solomon (12,9) on full link control words -----------------------------
def _lc_script(seed, variant):
    """Burst descriptors of slot 0 (slot 1 idles): calls of header / two voice superframes / terminator.  Returns
    (bursts as sent, bursts a decoder WITHOUT the RS check must see to behave like one WITH it)."""
    rng = np.random.default_rng(seed)
    sent, equiv_verify, equiv_correct = [], [], []
    for call in range(6):
        lc9 = synth.dmr_full_lc(0 if call % 2 else 3, int(rng.integers(1, 1 << 24)), int(rng.integers(1, 1 << 24)))[:9]
        for data_type, mask in ((synth.DMR_DT_VOICE_LC, 0x96), (synth.DMR_DT_TERMINATOR_LC, 0x99)):
            good = lc9 + synth.dmr_rs_12_9_parity(lc9, mask)
            if variant == "valid":
                tx, v, c = good, ("data", data_type, good), ("data", data_type, good)
            elif variant == "random":
                bad = lc9 + [int(t) ^ 0x5A for t in good[9:]]
                tx, v, c = bad, ("data", synth.DMR_DT_CSBK, bad), ("data", synth.DMR_DT_CSBK, bad)
            else:   # one octet of the 12 is wrong: detected by "verify", repaired by "correct"
                pos = int(rng.integers(0, 12))
                bad = list(good)
                bad[pos] ^= int(rng.integers(1, 256))
                tx, v, c = bad, ("data", synth.DMR_DT_CSBK, bad), ("data", data_type, good)
            if data_type == synth.DMR_DT_VOICE_LC:
                head = (("data", data_type, tx), v, c)
            else:
                tail = (("data", data_type, tx), v, c)
        voice = [("voice", None, None)] + [("voice", (1, 0, 0), None)] * 5
        for lst, k in ((sent, 0), (equiv_verify, 1), (equiv_correct, 2)):
            lst.append(head[k])
            lst.extend(voice * 2)
            lst.append(tail[k])
            lst.extend([("data", synth.DMR_DT_IDLE, [0] * 12)] * 2)
    return sent, equiv_verify, equiv_correct


def _render(script, seed):
    rng = np.random.default_rng(seed)
    out = []
    idle = ("data", synth.DMR_DT_IDLE, [0] * 12)
    for kind, a, b in script:
        for slot, (k, x, y) in enumerate(((kind, a, b), idle)):
            out.append(synth.dmr_data_burst(slot, x, y, rng) if k == "data" else synth.dmr_voice_burst(slot, rng, emb=x, fragment=y))
    return np.concatenate(out).astype(np.uint8)


@pytest.mark.parametrize("variant", ["valid", "random", "one_error"])
def test_dmr_opt_in_rs_12_9_on_full_lc(variant):
    """DH_OPT_DMR_LC_FEC (off by default: everything above runs with the reference's behaviour).  The reference has
    no RS check to compare with, so the expectation comes from the reference itself on an EQUIVALENT stream: a full-LC
    burst the check rejects must act like a burst type the decoder ignores (CSBK), a repaired one like the clean one."""
    import digiham_b200 as dh
    sent, eq_verify, eq_correct = _lc_script(7, variant)
    s_sent, s_verify, s_correct = _render(sent, 99), _render(eq_verify, 99), _render(eq_correct, 99)
    orc = oracle_lib.best()
    want = {0: orc.decode(oracle_lib.PROTO_DMR, s_sent), 1: orc.decode(oracle_lib.PROTO_DMR, s_verify),
            2: orc.decode(oracle_lib.PROTO_DMR, s_correct)}
    if variant != "valid":
        assert want[0][1] != want[1][1], "the two streams must differ in metadata for the test to mean anything"
    got = {}
    for mode in (0, 1, 2):
        bank = dh.DecoderBank(2, dh.PROTO_DMR)
        bank.set_option(dh.OPT_DMR_LC_FEC, mode, channel=0)      # channel 1 keeps the default
        sym = np.stack([s_sent, s_sent])
        out, meta = _gpu_decode(bank, sym, [5000, sym.shape[1] - 5000])
        assert out[0] == want[mode][0].tobytes() and meta[0] == want[mode][1], (variant, mode)
        assert out[1] == want[0][0].tobytes() and meta[1] == want[0][1], (variant, mode)
        got[mode] = meta[0]
        bank.close()
    assert b"source:" in got[0] and len(want[0][0]) > 27 * 20
    if variant == "random":
        assert b"source:" not in got[1] and b"source:" not in got[2]
    if variant == "one_error":
        assert got[1] != got[2] and b"source:" in got[2]
    ysf = dh.DecoderBank(1, dh.PROTO_YSF)
    with pytest.raises(dh.DhError):
        ysf.set_option(dh.OPT_DMR_LC_FEC, 1)
    ysf.close()
