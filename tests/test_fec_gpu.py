"""Device-level parity of the FEC primitives: the DEVICE block-code decoders, the BPTC(196,96)
column/row pivot and the register-exchange Viterbi decoders are driven directly (dh_test_* hooks of the C ABI) and
compared with the reference's own C functions, element by element:

  * Hamming(7,4) / (13,9) / (15,11) / (16,11), QR(16,7), Golay(20,8): ALL 2^n words;
  * Golay(24,12), BCH(31,21): every error pattern of weight <= 3 (<= 4 for Golay) on a set of codewords, plus 2^20
    random words;
  * BPTC(196,96): encoded payloads with 0..4 bit errors and random payloads;
  * Viterbi: 10^5 inputs per length — encoded data with few errors, with many errors, and uniformly random dibits
    (the highest path metrics the decoder can see).  The references keep the metric in a uint8_t (YSF,
    src/ysf_decoder/trellis.c:28,68) / uint16_t (NXDN); the device metric must equal it, and the test records the
    largest metric seen (it stays far below 256: the wrap cannot be reached with <= 180 steps).
"""
import ctypes
import itertools
import os

import numpy as np
import pytest

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 8


def _dev_fec(code, words):
    import digiham_b200 as dh
    w = np.ascontiguousarray(words, dtype=np.uint32).copy()
    ok = np.zeros(w.size, dtype=np.uint8)
    dh._capi.check(dh.lib().dh_test_fec(code, w.ctypes.data, ok.ctypes.data, w.size))
    return ok, w


def _compare_fec(code, words):
    orc = oracle_lib.best()
    ok_d, w_d = _dev_fec(code, words)
    ok_r, w_r = orc.fec_batch(code, words, threads=THREADS)
    bad = np.nonzero((ok_d != ok_r) | (w_d != w_r))[0]
    assert bad.size == 0, "%s: %d of %d words differ, first 0x%x: device (%d, 0x%x) reference (%d, 0x%x)" % (
        oracle_lib.FEC_NAMES[code], bad.size, len(words), int(words[bad[0]]), ok_d[bad[0]], w_d[bad[0]], ok_r[bad[0]],
        w_r[bad[0]])
    return ok_r


@pytest.mark.parametrize("code", [0, 1, 2, 3, 4, 5])
def test_block_codes_all_words(code):
    n = oracle_lib.FEC_BITS[code]
    ok = _compare_fec(code, np.arange(1 << n, dtype=np.uint32))
    assert ok.any()
    if code in (1, 4, 5):          # codes with unmapped syndromes must also report failures
        assert not ok.all()


def _patterns(nbits, max_weight):
    pats = [0]
    for wgt in range(1, max_weight + 1):
        for pos in itertools.combinations(range(nbits), wgt):
            v = 0
            for p in pos:
                v |= 1 << p
            pats.append(v)
    return np.array(pats, dtype=np.uint64)


def test_golay_24_12_patterns_and_random_words():
    rng = np.random.default_rng(1)
    cws = np.array([synth.encode_block("golay_24_12", int(d)) for d in [0, 0xFFF, 0x5A5] + list(rng.integers(0, 4096, 13))],
                   dtype=np.uint64)
    pats = _patterns(24, 4)                       # 1 + 24 + 276 + 2024 + 10626 patterns
    words = (cws[:, None] ^ pats[None, :]).reshape(-1).astype(np.uint32)
    ok = _compare_fec(6, words)
    w3 = (cws[:, None] ^ _patterns(24, 3)[None, :]).reshape(-1).astype(np.uint32)
    ok3, fixed = _dev_fec(6, w3)
    assert ok3.all() and np.array_equal(fixed.reshape(len(cws), -1), np.repeat(cws[:, None], 2325, axis=1).astype(np.uint32))
    assert not ok.all()                           # weight-4 patterns are detected, not corrected
    _compare_fec(6, rng.integers(0, 1 << 24, size=1 << 20, dtype=np.uint32))


def test_bch_31_21_patterns_and_random_words():
    rng = np.random.default_rng(2)
    cws = np.array([synth.bch_31_21_encode(int(d)) for d in [0, (1 << 21) - 1] + list(rng.integers(0, 1 << 21, 14))],
                   dtype=np.uint64)
    pats = _patterns(31, 3)                       # 1 + 31 + 465 + 4495
    words = (cws[:, None] ^ pats[None, :]).reshape(-1).astype(np.uint32)
    _compare_fec(7, words)
    w2 = (cws[:, None] ^ _patterns(31, 2)[None, :]).reshape(-1).astype(np.uint32)
    ok2, fixed = _dev_fec(7, w2)
    assert ok2.all() and np.array_equal(fixed.reshape(len(cws), -1), np.repeat(cws[:, None], 497, axis=1).astype(np.uint32))
    _compare_fec(7, rng.integers(0, 1 << 31, size=1 << 20, dtype=np.uint32))


def test_bptc_196_96_device_vs_reference():
    import digiham_b200 as dh
    rng = np.random.default_rng(3)
    n_enc, n_rand = 30000, 10000
    payloads = np.zeros((n_enc + n_rand, 25), dtype=np.uint8)
    infos = rng.integers(0, 256, size=(n_enc, 12), dtype=np.uint8)
    for i in range(n_enc):
        if i < 2000 or i % 15 == 0:
            bits = synth.dmr_bptc_encode(infos[i])
        else:
            bits = base.copy()
        base = bits
        nerr = i % 5
        flip = rng.choice(196, size=nerr, replace=False)
        b = bits.copy()
        b[flip] ^= 1
        payloads[i] = np.packbits(np.concatenate([b, np.zeros(4, dtype=np.uint8)]))
    payloads[n_enc:] = rng.integers(0, 256, size=(n_rand, 25), dtype=np.uint8)
    out = np.zeros((payloads.shape[0], 12), dtype=np.uint8)
    ok = np.zeros(payloads.shape[0], dtype=np.uint8)
    dh._capi.check(dh.lib().dh_test_bptc(payloads.ctypes.data, out.ctypes.data, ok.ctypes.data, payloads.shape[0]))
    ok_r, out_r = oracle_lib.best().bptc_batch(payloads, threads=THREADS)
    assert np.array_equal(ok, ok_r)
    good = ok_r.astype(bool)
    assert np.array_equal(out[good], out_r[good])
    assert good[:n_enc].mean() > 0.7 and not good.all()
    # error-free encodings decode to their info bytes
    clean = np.arange(0, 2000, 5)
    assert good[clean].all() and np.array_equal(out[clean], infos[clean])


def _trellis_inputs(steps, n, rng):
    """A third encoded data with 0-3 symbol errors, a third with ~15 % symbol errors, a third uniformly random."""
    d = np.zeros((n, steps), dtype=np.uint8)
    third = n // 3
    data = rng.integers(0, 2, size=(2 * third, steps), dtype=np.uint8)
    data[:, -4:] = 0
    pool = [synth.ysf_conv_encode(data[i]) for i in range(min(2 * third, 3000))]
    for i in range(2 * third):
        d[i] = pool[i % len(pool)]
    few = rng.integers(0, 4, size=third)
    for i in range(third):
        pos = rng.choice(steps, size=few[i], replace=False)
        d[i, pos] ^= rng.integers(1, 4, size=few[i]).astype(np.uint8)
    noisy = rng.random((third, steps)) < 0.15
    d[third:2 * third] ^= (noisy * rng.integers(1, 4, size=(third, steps))).astype(np.uint8)
    d[2 * third:] = rng.integers(0, 4, size=(n - 2 * third, steps), dtype=np.uint8)
    return d


@pytest.mark.parametrize("variant,steps,nxdn", [(0, 100, False), (1, 180, False), (2, 36, True), (3, 96, True)])
def test_viterbi_device_vs_reference(variant, steps, nxdn):
    import digiham_b200 as dh
    rng = np.random.default_rng(10 + variant)
    n = 100001                                    # odd: the paired YSF kernel also sees a lone last input
    d = _trellis_inputs(steps, n, rng)
    nw = (steps + 31) // 32
    words = np.zeros((n, nw), dtype=np.uint32)
    metric = np.zeros(n, dtype=np.uint32)
    dh._capi.check(dh.lib().dh_test_viterbi(variant, d.ctypes.data, n, words.ctypes.data, metric.ctypes.data))
    m_r, out_r = oracle_lib.best().trellis_batch(d, nxdn=nxdn, threads=THREADS)
    nbytes = (steps + 7) // 8
    got = words.byteswap().view(np.uint8).reshape(n, nw * 4)[:, :nbytes]
    if nxdn:
        # Nxdn::Trellis::decode writes (len + 15) / 16 bytes: compare the bytes both produce
        nbytes = (2 * steps + 15) // 16
    assert np.array_equal(got[:, :nbytes], out_r[:, :nbytes])
    assert np.array_equal(metric, m_r)
    assert m_r[:n // 3].max() <= 6 and m_r.max() < 200, "largest path metric %d" % m_r.max()
    assert m_r[2 * (n // 3):].mean() > 0.1 * steps  # the random third really is a high-error workload
