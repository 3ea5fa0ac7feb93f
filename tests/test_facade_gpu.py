"""BASELINE config 1 (plumbing): the header-compatible C++ facade modules of include/*.hpp, chained through csdr
ring buffers by a clone of the reference's CLI loop (tests/cpp/facade_pipe.cpp), produce exactly what the CPU
oracle's modules produce on the same one-channel input."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from digiham_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = oracle_lib.ROOT


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("facade")
    exe = str(d / "facade_pipe")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "oracle", "csdr_shim"), os.path.join(ROOT, "tests", "cpp", "facade_pipe.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "digiham_b200"), "-ldigiham_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "digiham_b200")], check=True)
    return exe, d


def _run(harness, proto, data):
    exe, d = harness
    inp = str(d / (proto + ".in"))
    data.tofile(inp)
    prefix = str(d / proto)
    for ext in (".out", ".meta"):
        if os.path.exists(prefix + ext):
            os.remove(prefix + ext)
    subprocess.run([exe, proto, inp, prefix], check=True, stderr=subprocess.PIPE)
    out = np.fromfile(prefix + ".out", dtype=np.uint8)
    meta = open(prefix + ".meta", "rb").read() if os.path.exists(prefix + ".meta") else b""
    return out, meta


def test_facade_dmr_pipe(harness):
    x, _ = synth.dmr_channel_bank(1, 60000, seed=31, device="cpu", noise_fraction=0.0)
    x = x[0, :60000].numpy()
    out, meta = _run(harness, "dmr", x)
    _, ref_out, ref_meta = oracle_lib.best().pipe(oracle_lib.PROTO_DMR, x, chunk=128)
    assert np.array_equal(out, ref_out) and meta == ref_meta
    assert out.size >= 27 and b"protocol:DMR" in meta


def test_facade_ysf_pipe(harness):
    sym = synth.ysf_symbols(12, seed=8, mode="DN", lead_in=50)
    x = synth.modulate(sym, sps=10, snr_db=18, rng=np.random.default_rng(1), phase=4)
    out, meta = _run(harness, "ysf", x)
    _, ref_out, ref_meta = oracle_lib.best().pipe(oracle_lib.PROTO_YSF, x, chunk=128)
    assert np.array_equal(out, ref_out) and meta == ref_meta
    assert out.size > 0 and b"mode:DN" in meta


def test_facade_nxdn_pipe(harness):
    sym = synth.nxdn_symbols(40, seed=8, lead_in=50)
    x = synth.modulate(sym, sps=20, snr_db=18, rng=np.random.default_rng(1), phase=4)
    out, meta = _run(harness, "nxdn", x)
    _, ref_out, ref_meta = oracle_lib.best().pipe(oracle_lib.PROTO_NXDN, x, chunk=128)
    assert np.array_equal(out, ref_out) and meta == ref_meta
    assert out.size > 0 and b"protocol:NXDN;sync:voice" in meta


def test_facade_dstar_pipe(harness):
    sym = np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), 150), synth.dstar_symbols(80, seed=8, lead_in=0)])
    x = synth.modulate(sym, sps=10, levels=synth.LEVELS2, snr_db=18, rng=np.random.default_rng(1), phase=4)
    out, meta = _run(harness, "dstar", x)
    _, ref_out, ref_meta = oracle_lib.best().pipe(oracle_lib.PROTO_DSTAR, x, chunk=128)
    assert np.array_equal(out, ref_out) and meta == ref_meta
    assert out.size > 0 and b"protocol:DSTAR" in meta


def test_facade_pocsag_pipe(harness):
    bits = synth.pocsag_bits([(1234562, 3, "HELLO B200"), (42, 3, "FACADE")], seed=2, bit_errors=1)
    x = synth.modulate(bits, sps=40, levels=synth.LEVELS2[::-1].copy(), snr_db=20, rng=np.random.default_rng(3))
    out, meta = _run(harness, "pocsag", x)
    _, ref_out, _ = oracle_lib.best().pipe(oracle_lib.PROTO_POCSAG, x, chunk=128)
    assert np.array_equal(out, ref_out) and meta == b""
    assert b"message:HELLO B200" in out.tobytes()


def test_facade_pocsag_custom_serializer_gets_structured_records(harness):
    """Pocsag::Decoder(Serializer*) (reference include/pocsag_decoder.hpp:12, src/pocsag_decoder/message.cpp:16-24):
    a caller-supplied serializer sees the {address, message} map itself — message bodies that contain the text
    separators of the default rendering (';message:', a newline followed by 'address:') arrive intact."""
    tricky = ["PLAIN", "A;message:B", "X\naddress:7;message:Y", "tail:;\n"]
    msgs = [(1000 + 8 * k, 3, t) for k, t in enumerate(tricky)]
    bits = synth.pocsag_bits(msgs, seed=5, lead_in=0)
    x = synth.modulate(bits, sps=40, levels=synth.LEVELS2[::-1].copy(), snr_db=25, rng=np.random.default_rng(3))
    out, _ = _run(harness, "pocsag_custom", x)
    want = b"".join(("address=%d:%d" % (len(str(a)), a)).encode() + ("message=%d:" % len(t)).encode() + t.encode() + b"\x1e"
                    for a, _, t in msgs)
    assert out.tobytes() == want
    # the default serializer on the same signal still equals the reference byte for byte
    out, _ = _run(harness, "pocsag", x)
    _, ref_out, _ = oracle_lib.best().pipe(oracle_lib.PROTO_POCSAG, x, chunk=128)
    assert np.array_equal(out, ref_out)


def test_facade_rrc_and_dvf(harness):
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, 5000).astype(np.float32)
    out, _ = _run(harness, "rrc", x)
    assert np.array_equal(out.view(np.uint32), oracle_lib.best().rrc(x, chunk=128).view(np.uint32))
    a = rng.integers(-20000, 20000, 4000).astype(np.int16)
    out, _ = _run(harness, "dvf", a)
    assert np.array_equal(out.view(np.int16), oracle_lib.best().dvf(a, chunk=128))


def test_c_abi_example_program(tmp_path):
    """examples/many_channels.cpp: the streaming C-ABI call sequence of INTEGRATION.md compiles, runs and decodes."""
    exe = str(tmp_path / "many_channels")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "examples", "many_channels.cpp"), "-L" + os.path.join(ROOT, "digiham_b200"),
                    "-ldigiham_b200", "-Wl,-rpath," + os.path.join(ROOT, "digiham_b200"), "-L/usr/local/cuda/lib64",
                    "-lcudart", "-o", exe], check=True)
    r = subprocess.run([exe, "64", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "64 channels x 3 steps" in r.stdout
