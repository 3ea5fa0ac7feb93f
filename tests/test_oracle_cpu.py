"""CPU suite, part 1: the oracle(s) against the committed golden fixtures (tests/golden/golden_v1.npz, generated from
the compiled reference by tests/golden/make_golden.py), and the two oracles against each other where both exist.
Nothing here needs a GPU."""
import os

import numpy as np
import pytest

import oracle_lib

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
GOLDEN2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v2.npz"))   # NXDN / D-Star


def _oracles():
    out = []
    if oracle_lib.ref() is not None:
        out.append(oracle_lib.ref())
    if os.listdir(os.path.join(oracle_lib.ORACLE_DIR, "port")):
        out.append(oracle_lib.port())
    return out


ORACLES = _oracles()
IDS = [o.kind for o in ORACLES]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_an_oracle_exists():
    assert ORACLES, "neither oracle/_ref/libdigiham_ref.so nor oracle/port is available"


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_rrc_golden(orc):
    x = GOLDEN["rrc_in"]
    assert np.array_equal(_bits(orc.rrc(x, narrow=False)), _bits(GOLDEN["rrc_wide_out"]))
    assert np.array_equal(_bits(orc.rrc(x, narrow=True)), _bits(GOLDEN["rrc_narrow_out"]))
    # chunked feeding changes nothing (src/lib/cli.cpp:29-33 drain loop)
    assert np.array_equal(_bits(orc.rrc(x, narrow=False, chunk=37)), _bits(GOLDEN["rrc_wide_out"]))


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_demod_golden(orc):
    assert np.array_equal(orc.demod(GOLDEN["gfsk10_in"], sps=10, four_level=True), GOLDEN["gfsk10_out"])
    assert np.array_equal(orc.demod(GOLDEN["gfsk10_in"], sps=10, four_level=True, chunk=128), GOLDEN["gfsk10_out"])
    assert np.array_equal(orc.demod(GOLDEN["fsk40_in"], sps=40, four_level=False, invert=True),
                          GOLDEN["fsk40_inv_out"])
    assert np.array_equal(orc.demod(GOLDEN["fsk40_in"], sps=40, four_level=False, invert=False), GOLDEN["fsk40_out"])
    # the inverted output is the complement
    assert np.array_equal(GOLDEN["fsk40_inv_out"] ^ 1, GOLDEN["fsk40_out"])


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_dmr_decoder_golden(orc):
    for k in range(2):
        out, meta = orc.decode(oracle_lib.PROTO_DMR, GOLDEN["dmr%d_sym" % k])
        assert np.array_equal(out, GOLDEN["dmr%d_out" % k])
        assert meta == GOLDEN["dmr%d_meta" % k].tobytes()
        out2, meta2 = orc.decode(oracle_lib.PROTO_DMR, GOLDEN["dmr%d_sym" % k], chunk=100)
        assert np.array_equal(out2, out) and meta2 == meta
    assert b"sync:voice" in GOLDEN["dmr0_meta"].tobytes() and b"talkeralias:B200 TESTER" in GOLDEN["dmr0_meta"].tobytes()
    assert GOLDEN["dmr0_out"].size % 27 == 0 and GOLDEN["dmr0_out"].size > 0


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_pipe_golden(orc):
    for c in range(2):
        sym, out, meta = orc.pipe(oracle_lib.PROTO_DMR, GOLDEN["pipe%d_in" % c], chunk=1000)
        assert np.array_equal(sym, GOLDEN["pipe%d_sym" % c])
        assert np.array_equal(out, GOLDEN["pipe%d_out" % c])
        assert meta == GOLDEN["pipe%d_meta" % c].tobytes()


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_block_codes_golden(orc):
    for cid, name in enumerate(oracle_lib.FEC_NAMES):
        table = GOLDEN["fec_" + name]
        for s in range(table.shape[0]):
            ok, w = orc.fec(cid, s)
            assert (int(ok), w) == (int(table[s, 0]), int(table[s, 1])), (name, s)
    for p, exp in zip(GOLDEN["bptc_in"], GOLDEN["bptc_out"]):
        ok, out = orc.bptc(p)
        assert int(ok) == exp[0]
        if ok:
            assert np.array_equal(out, exp[1:])
    assert GOLDEN["bptc_out"][:10, 0].all(), "valid BPTC blocks with <= 4 bit errors must decode"


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_ysf_pocsag_dvf_golden(orc):
    for k in range(3):
        out, meta = orc.decode(oracle_lib.PROTO_YSF, GOLDEN["ysf%d_sym" % k])
        assert np.array_equal(out, GOLDEN["ysf%d_out" % k]), k
        assert meta == GOLDEN["ysf%d_meta" % k].tobytes(), k
    assert b"mode:DN" in GOLDEN["ysf0_meta"].tobytes() and b"lat:" in GOLDEN["ysf0_meta"].tobytes()
    out, _ = orc.decode(oracle_lib.PROTO_POCSAG, GOLDEN["pocsag_bits"])
    assert np.array_equal(out, GOLDEN["pocsag_out"]) and b"message:THE QUICK BROWN FOX" in out.tobytes()
    sym, out, _ = orc.pipe(oracle_lib.PROTO_POCSAG, GOLDEN["pocsag_pipe_in"])
    assert np.array_equal(sym, GOLDEN["pocsag_pipe_sym"]) and np.array_equal(out, GOLDEN["pocsag_pipe_out"])
    _, out, meta = orc.pipe(oracle_lib.PROTO_YSF, GOLDEN["ysf_pipe_in"])
    assert np.array_equal(out, GOLDEN["ysf_pipe_out"]) and meta == GOLDEN["ysf_pipe_meta"].tobytes()
    assert np.array_equal(orc.dvf(GOLDEN["dvf_in"]), GOLDEN["dvf_out"])
    assert np.array_equal(orc.dvf(GOLDEN["dvf_in"], chunk=77), GOLDEN["dvf_out"])


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_ysf_primitives_golden(orc):
    for steps in (100, 180):
        for packed, exp in zip(GOLDEN["trellis%d_in" % steps], GOLDEN["trellis%d_out" % steps]):
            metric, bits = orc.trellis(packed, steps)
            assert metric == exp[0] and np.array_equal(bits, exp[1:])
        assert GOLDEN["trellis%d_out" % steps][0, 0] == 0      # an error-free code sequence decodes with metric 0
    blob = GOLDEN["crc_in"]
    assert [orc.crc16(blob[:k]) for k in (4, 10, 20, 40)] == [int(v) for v in GOLDEN["crc_out"]]
    assert np.array_equal(orc.whitening(blob[:20], 160), GOLDEN["whitening_out"])


def test_generated_luts_match_reference_for_every_syndrome():
    """tools/gen_tables.py derives H and the correction LUTs from the published generator matrices; the result
    must equal what the reference's linear search over corrections[] does (golden = compiled reference)."""
    import re
    src = open(os.path.join(oracle_lib.ROOT, "digiham_b200", "csrc", "tables.inc")).read()

    def macro(name):
        m = re.search(r"#define %s \{ \\\n(.*?)\n\}" % name, src, re.S)
        return [int(v.strip().rstrip("u"), 16) for v in m.group(1).replace("\\", " ").replace("\n", " ").split(",")
                if v.strip()]

    for name in oracle_lib.FEC_NAMES:
        lut = macro("DH_%s_LUT_INIT" % name.upper())
        table = GOLDEN["fec_" + name]
        assert len(lut) == table.shape[0]
        for s in range(len(lut)):
            ok = s == 0 or lut[s] != 0
            assert ok == bool(table[s, 0]), (name, s)
            if ok:
                assert (s ^ lut[s]) == int(table[s, 1]), (name, s)


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_nxdn_golden(orc):
    for k in range(3):
        out, meta = orc.decode(oracle_lib.PROTO_NXDN, GOLDEN2["nxdn%d_sym" % k])
        assert np.array_equal(out, GOLDEN2["nxdn%d_out" % k]), k
        assert meta == GOLDEN2["nxdn%d_meta" % k].tobytes(), k
    all_meta = b"".join(GOLDEN2["nxdn%d_meta" % k].tobytes() for k in range(3))
    for key in (b"sync:voice", b"source:", b"destination:", b"type:"):
        assert key in all_meta, key
    sym, out, meta = orc.pipe(oracle_lib.PROTO_NXDN, GOLDEN2["nxdn_pipe_in"])
    assert np.array_equal(sym, GOLDEN2["nxdn_pipe_sym"]) and np.array_equal(out, GOLDEN2["nxdn_pipe_out"])
    assert meta == GOLDEN2["nxdn_pipe_meta"].tobytes() and out.size >= 18
    for nbits in (72, 192):
        for row, want in zip(GOLDEN2["nxdn_trellis%d_in" % nbits], GOLDEN2["nxdn_trellis%d_out" % nbits]):
            metric, bits = orc.nxdn_trellis(row, nbits)
            assert metric == want[0] and np.array_equal(bits, want[1:])
    for d, want in zip(GOLDEN2["nxdn_sacch_in"], GOLDEN2["nxdn_sacch_out"]):
        ok, data = orc.nxdn_sacch(d)
        assert int(ok) == want[0] and (not ok or np.array_equal(data, want[1:]))
    assert [orc.nxdn_facch1(d) for d in GOLDEN2["nxdn_facch1_in"]] == list(GOLDEN2["nxdn_facch1_out"])
    assert 0x08 in GOLDEN2["nxdn_facch1_out"] and -1 in GOLDEN2["nxdn_facch1_out"]


@pytest.mark.parametrize("orc", ORACLES, ids=IDS)
def test_dstar_golden(orc):
    for k in range(3):
        out, meta = orc.decode(oracle_lib.PROTO_DSTAR, GOLDEN2["dstar%d_sym" % k])
        assert np.array_equal(out, GOLDEN2["dstar%d_out" % k]), k
        assert meta == GOLDEN2["dstar%d_meta" % k].tobytes(), k
    all_meta = b"".join(GOLDEN2["dstar%d_meta" % k].tobytes() for k in range(3))
    for key in (b"ourcall:", b"message:", b"lat:", b"dprs:"):
        assert key in all_meta, key
    sym, out, meta = orc.pipe(oracle_lib.PROTO_DSTAR, GOLDEN2["dstar_pipe_in"])
    assert np.array_equal(sym, GOLDEN2["dstar_pipe_sym"]) and np.array_equal(out, GOLDEN2["dstar_pipe_out"])
    assert meta == GOLDEN2["dstar_pipe_meta"].tobytes() and out.size >= 9
    for bits, want in zip(GOLDEN2["dstar_header_in"], GOLDEN2["dstar_header_out"]):
        rc, text = orc.dstar_header(bits)
        assert ("%d|" % rc).encode() + text == want.tobytes().rstrip(b"\0")


@pytest.mark.skipif(len(ORACLES) < 2, reason="needs both the compiled reference and the port")
def test_port_equals_reference_nxdn_dstar():
    ref, port = oracle_lib.ref(), oracle_lib.port()
    from digiham_b200 import synth
    rng = np.random.default_rng(6)
    for k in range(6):
        sym = synth.nxdn_symbols(80, seed=60 + k, symbol_errors=[0.0, 0.01, 0.05][k % 3])
        a, b = ref.decode(oracle_lib.PROTO_NXDN, sym), port.decode(oracle_lib.PROTO_NXDN, sym)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1], k
        bits = synth.dstar_symbols(250, seed=70 + k, bit_errors=[0.0, 0.003, 0.03][k % 3])
        a, b = ref.decode(oracle_lib.PROTO_DSTAR, bits), port.decode(oracle_lib.PROTO_DSTAR, bits)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1], k
    noise4 = rng.integers(0, 4, 60000).astype(np.uint8)
    for proto in (oracle_lib.PROTO_NXDN, oracle_lib.PROTO_DSTAR):
        a, b = ref.decode(proto, noise4), port.decode(proto, noise4)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    for t in range(40):
        nb = (72, 192)[t % 2]
        d = rng.integers(0, 256, nb // 8).astype(np.uint8)
        a, b = ref.nxdn_trellis(d, nb), port.nxdn_trellis(d, nb)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
    x = synth.modulate(synth.nxdn_symbols(30, seed=77), sps=20, snr_db=14, ppm=40, rng=np.random.default_rng(8))
    a, b = ref.pipe(oracle_lib.PROTO_NXDN, x), port.pipe(oracle_lib.PROTO_NXDN, x)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


@pytest.mark.skipif(len(ORACLES) < 2, reason="needs both the compiled reference and the port")
def test_port_equals_reference_on_random_streams():
    ref, port = oracle_lib.ref(), oracle_lib.port()
    from digiham_b200 import synth
    rng = np.random.default_rng(5)
    x = rng.normal(0, 0.3, 5000).astype(np.float32)
    for narrow in (False, True):
        assert np.array_equal(_bits(ref.rrc(x, narrow)), _bits(port.rrc(x, narrow)))
    xb, _ = synth.dmr_channel_bank(6, 40000, seed=9, device="cpu")
    for c in range(6):
        a = ref.pipe(oracle_lib.PROTO_DMR, xb[c, :40000].numpy())
        b = port.pipe(oracle_lib.PROTO_DMR, xb[c, :40000].numpy())
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    for k, mode in enumerate(["DN", "V1", "VW", "mix", "FR"]):
        sym = synth.ysf_symbols(25, seed=50 + k, mode=mode, symbol_errors=[0.0, 0.01, 0.04][k % 3])
        a = ref.decode(oracle_lib.PROTO_YSF, sym)
        b = port.decode(oracle_lib.PROTO_YSF, sym)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1], mode
    for k in range(4):
        bits = synth.pocsag_bits([(100 + k, 3, "PORT VS REFERENCE %d" % k), (5, 1, ""), (77777, 3, "x" * 70)],
                                 seed=k, bit_errors=k, lead_in=17 * k)
        assert np.array_equal(ref.decode(oracle_lib.PROTO_POCSAG, bits)[0], port.decode(oracle_lib.PROTO_POCSAG, bits)[0])
    a16 = rng.integers(-32768, 32768, 6000).astype(np.int16)
    assert np.array_equal(ref.dvf(a16), port.dvf(a16))
    for sps, four, inv in ((10, True, False), (20, True, False), (40, False, True), (7, False, False)):
        sig = synth.modulate(synth.random_symbols(700, 4 if four else 2, seed=sps),
                             sps=sps, levels=synth.LEVELS4 if four else synth.LEVELS2, ppm=500, snr_db=12,
                             rng=np.random.default_rng(sps))
        assert np.array_equal(ref.demod(sig, sps, four, inv), port.demod(sig, sps, four, inv)), sps
