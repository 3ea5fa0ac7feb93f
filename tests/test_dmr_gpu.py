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
