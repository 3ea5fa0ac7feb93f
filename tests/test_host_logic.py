"""CPU suite, part 2: host-side logic of the product and the C-ABI surface.  No GPU, no compute calls."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import digiham_b200
    return digiham_b200.lib()


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "digiham_b200.h")).read()
    declared = sorted(set(re.findall(r"DH_API[^;(]*?\b(dh_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = _lib()
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, "declared in include/digiham_b200.h but not exported: %s" % missing
    nm = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "digiham_b200", "libdigiham_b200.so")],
                        stdout=subprocess.PIPE, text=True).stdout
    exported = set(re.findall(r" T (dh_[a-z0-9_]+)", nm))
    assert exported == set(declared), "exported but undeclared: %s" % sorted(exported - set(declared))
    assert L.dh_version().decode().startswith("0.")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib()
    h = ctypes.c_void_p()
    for rc in (L.dh_rrc_create(ctypes.byref(h), 0, 4, 0),
               L.dh_demod_create(ctypes.byref(h), 0, 4, 1, 10, 0),
               L.dh_decoder_create(ctypes.byref(h), 0, 4, 0),
               L.dh_pipe_create(ctypes.byref(h), 0, 4, 0, 1000)):
        assert rc != 0 and not h.value
    assert b"no CUDA device" in L.dh_last_error()
    import digiham_b200 as dh
    with pytest.raises(dh.DhError):
        dh.RrcBank(4)


def test_product_does_not_reference_the_oracle():
    bad = []
    for base in ("digiham_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".cu", ".cuh", ".hpp", ".h", ".py", ".cpp", ".inc")) and "build" not in dp:
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"oracle_api|liboracle|libdigiham_ref|oracle_lib|/oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


RECIP_SRC = r"""
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
int main(void) {
    const double gains[2] = {8.337797030e+00, 1.667711971e+01};
    unsigned long long bad = 0;
    for (int g = 0; g < 2; g++) {
        const double gain = gains[g], rg = 1.0 / gain;
        #pragma omp parallel for reduction(+:bad) schedule(static)
        for (long long b = 0; b < (1LL << 32); b += STRIDE) {
            uint32_t u = (uint32_t) b, r1, r2; float s, q1, q2;
            memcpy(&s, &u, 4);
            if (isnan(s)) continue;
            q1 = (float) ((double) s / gain);      /* reference: src/rrc_filter/rrc_filter.cpp:33 */
            q2 = (float) ((double) s * rg);        /* K1 epilogue */
            memcpy(&r1, &q1, 4); memcpy(&r2, &q2, 4);
            if (r1 != r2) bad++;
        }
    }
    printf("%llu\n", bad);
    return 0;
}
"""


def test_reciprocal_gain_exhaustive():
    """K1 scales by fl64(1/gain) instead of dividing; for the two built-in gains this is bit-identical for EVERY
    float32 input (all 2^32 patterns are tried; set DH_FAST_TESTS=1 to sample every 7th)."""
    stride = 7 if os.environ.get("DH_FAST_TESTS") else 1
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "recip.c")
        open(src, "w").write(RECIP_SRC)
        exe = os.path.join(d, "recip")
        subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-DSTRIDE=%d" % stride, src, "-o", exe, "-lm"],
                       check=True)
        out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert int(out.strip()) == 0


S16_SRC = r"""
#include <stdio.h>
#include <math.h>
int main(void) {
    int bad = 0;
    volatile float r = 1.0f / 32767.0f;
    for (int v = -32768; v < 32768; v++) {
        float f = (float) v;
        float ref = f / 32767.0f;                          /* csdr convert -i s16 -o float: (float) in / SHRT_MAX */
        float q = f * r;                                   /* K1's s16_to_float (rrc.cu) */
        float o = fmaf(fmaf(-q, 32767.0f, f), r, q);
        if (o != ref || signbit(o) != signbit(ref)) bad++;
    }
    printf("%d\n", bad);
    return 0;
}
"""


def test_s16_conversion_exhaustive():
    """The fused int16 -> float32 conversion of K1 (multiply by fl32(1/32767) + one fma refinement step) equals the
    IEEE division `(float) s / 32767.0f` for all 65536 inputs."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s16.c")
        open(src, "w").write(S16_SRC)
        exe = os.path.join(d, "s16")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", src, "-o", exe, "-lm"], check=True)
        out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert int(out.strip()) == 0


DIVC_SRC = r"""
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
int main(void) {
    const float cs[6] = {10.0f, 100.0f, 20.0f, 40.0f, 6.0f, 14.0f};   /* sps, 100, and hi - lo of the fast paths */
    unsigned long long bad = 0;
    for (int g = 0; g < 6; g++) {
        const float c = cs[g];
        const double rc = 1.0 / (double) c;
        #pragma omp parallel for reduction(+:bad) schedule(static)
        for (long long b = 0; b < (1LL << 32); b += STRIDE) {
            uint32_t u = (uint32_t) b, r1, r2; float x, q1, q2;
            memcpy(&x, &u, 4);
            if (isnan(x)) continue;
            q1 = x / c;                            /* reference: volume_sum / samplesPerSymbol, total / 100 */
            q2 = (float) ((double) x * rc);        /* K2 fast path */
            memcpy(&r1, &q1, 4); memcpy(&r2, &q2, 4);
            if (r1 != r2) bad++;
        }
    }
    printf("%llu\n", bad);
    return 0;
}
"""


def test_division_by_constant_exhaustive():
    """K2's fast paths (sps 10, 20, 40) compute x / sps, x / (hi - lo) and x / 100 as fl32(fl64(x) * fl64(1/c)); identical to the float division
    of the reference (gfsk_demodulator.cpp:53,82) for EVERY float32 x (DH_FAST_TESTS=1 samples every 7th)."""
    stride = 7 if os.environ.get("DH_FAST_TESTS") else 1
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "divc.c")
        open(src, "w").write(DIVC_SRC)
        exe = os.path.join(d, "divc")
        subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-DSTRIDE=%d" % stride, src, "-o", exe, "-lm"],
                       check=True)
        out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert int(out.strip()) == 0


def _events(recs):
    buf = bytearray()
    for kind, slot, a, b, data in recs:
        d = bytes(data) + bytes(12 - len(data))
        buf += bytes([kind, slot, a, b]) + d
    return bytes(buf)


def _replay(recs, proto=0):
    L = _lib()
    L.dh_meta_replay.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t,
                                 ctypes.POINTER(ctypes.c_size_t)]
    ev = _events(recs)
    out = ctypes.create_string_buffer(1 << 16)
    n = ctypes.c_size_t()
    assert L.dh_meta_replay(proto, ev, len(recs), out, len(out), ctypes.byref(n)) == 0
    return out.raw[:n.value].decode()


def test_dmr_meta_replay_lines():
    """Host restatement of Dmr::Slot / MetaCollector / handleLc (reference src/dmr_decoder/dmr_meta.cpp:7-179,
    dmr_phase.cpp:304-339): dirty tracking, key order, value formats."""
    lc_group = [0, 0, 0, 0x00, 0xAB, 0xCD, 0x12, 0x34, 0x56]
    lines = _replay([(2, 0, 1, 0, []),            # slot 0 sync:data
                     (2, 0, 1, 0, []),            # unchanged -> no line
                     (4, 0, 0, 0, lc_group),      # group call
                     (2, 0, 2, 0, []),            # sync:voice
                     (2, 1, 2, 0, []),            # slot 1 sync:voice
                     (3, 0, 0, 0, []),            # soft reset keeps sync
                     (1, 1, 0, 0, []),            # reset slot 1: nothing left to print but the slot
                     (1, 1, 0, 0, [])]).split("\n")
    assert lines == ["protocol:DMR;slot:0;sync:data",
                     "protocol:DMR;slot:0;source:1193046;sync:data;target:43981;type:group",
                     "protocol:DMR;slot:0;source:1193046;sync:voice;target:43981;type:group",
                     "protocol:DMR;slot:1;sync:voice",
                     "protocol:DMR;slot:0;sync:voice",
                     "protocol:DMR;slot:1",
                     ""]


def test_dmr_meta_replay_talker_alias_and_gps():
    name = b"B200 TESTER"
    hdr = [4, 0, (1 << 6) | (len(name) << 1)] + list(name[:6])
    blk = [5, 0] + list(name[6:]) + [0] * (7 - len(name[6:]))
    gps = [8, 0, 0x01, 0x12, 0x34, 0x56, 0x23, 0x45, 0x67]
    text = _replay([(2, 1, 2, 0, []), (4, 1, 0, 0, hdr), (4, 1, 0, 0, blk), (4, 1, 0, 0, gps),
                    (5, 1, 0, 0, []), (4, 1, 0, 0, blk)])
    lines = text.split("\n")
    assert lines[0] == "protocol:DMR;slot:1;sync:voice"
    assert lines[1] == "protocol:DMR;slot:1;sync:voice;talkeralias:B200 TESTER"
    assert lines[2] == "lat:24.799994;lon:-12.799995;protocol:DMR;slot:1;sync:voice;talkeralias:B200 TESTER"
    assert lines[3] == ""     # after the collector reset a lone block 1 is incomplete -> no change


def test_nxdn_meta_replay_lines():
    """Host restatement of Nxdn::MetaCollector (reference src/nxdn_decoder/nxdn_meta.cpp:6-76): every setter sends
    on its own, reset() batches, zero ids are not printed."""
    lines = _replay([(1, 0, 0, 0, []),                          # setSync("voice")
                     (1, 0, 0, 0, []),                          # unchanged
                     (2, 0, 1, 0, [0x12, 0x34, 0xAB, 0xCD]),    # conference, source 4660, destination 43981
                     (2, 0, 2, 0, [0x12, 0x34, 0x00, 0x00]),    # individual, destination 0 disappears
                     (2, 0, 0, 0, [0x12, 0x34, 0x00, 0x00]),    # other call type: type removed
                     (3, 0, 0, 0, []),
                     (3, 0, 0, 0, [])], proto=3).split("\n")
    assert lines == ["protocol:NXDN;sync:voice",
                     "protocol:NXDN;sync:voice;type:conference",
                     "protocol:NXDN;source:4660;sync:voice;type:conference",
                     "destination:43981;protocol:NXDN;source:4660;sync:voice;type:conference",
                     "destination:43981;protocol:NXDN;source:4660;sync:voice;type:individual",
                     "protocol:NXDN;source:4660;sync:voice;type:individual",
                     "protocol:NXDN;source:4660;sync:voice",
                     "protocol:NXDN",
                     ""]


def test_dstar_meta_replay_lines():
    """Host half of DStar::VoicePhase + MetaCollector (reference src/dstar_decoder/dstar_phase.cpp:151-278,
    dstar_meta.cpp:5-130): header fields, 20-character message, header resend with CRC, DPRS and GGA sentences."""
    from digiham_b200 import synth
    hdr = bytes(synth.dstar_header_bytes(my="DL1ABC", suffix="B200", your="CQCQCQ"))
    recs = [(1, 0, k, 0, hdr[12 * k:12 * k + 12]) for k in range(4)] + [(2, 0, 0, 0, [])]
    recs += [(3, 0, 0, 0, b) for b in synth.dstar_slow_data_blocks(message="HELLO FROM B200")]
    recs += [(4, 0, 1, 0, [])]
    hdr2 = bytes(synth.dstar_header_bytes(my="W1AW", suffix="", your="DL1ABC"))
    recs += [(3, 0, 0, 0, b) for b in synth.dstar_slow_data_blocks(header41=hdr2)]
    recs += [(3, 0, 0, 0, b) for b in synth.dstar_slow_data_blocks(text=synth.dstar_gga(4807.038, 1131.0, False, True))]
    recs += [(3, 0, 0, 0, b) for b in synth.dstar_slow_data_blocks(text=synth.dstar_dprs("X>Y:test"))]
    recs += [(4, 0, 0, 0, []), (5, 0, 0, 0, [])]
    lines = _replay(recs, proto=4).split("\n")
    base = "departure:DB0XYZ B;destination:DB0XYZ G;"
    assert lines[0] == base + "ourcall:DL1ABC/B200;protocol:DSTAR;sync:voice;yourcall:CQCQCQ"
    assert lines[1] == base + "message:HELLO FROM B200     ;ourcall:DL1ABC/B200;protocol:DSTAR;sync:voice;yourcall:CQCQCQ"
    assert lines[2] == base + "message:HELLO FROM B200     ;ourcall:W1AW;protocol:DSTAR;sync:voice;yourcall:DL1ABC"
    assert lines[3] == base + ("lat:48.117302;lon:-11.516666;message:HELLO FROM B200     ;ourcall:W1AW;protocol:DSTAR;"
                               "sync:voice;yourcall:DL1ABC")
    assert lines[4] == base + ("dprs:X>Y:test;lat:48.117302;lon:-11.516666;message:HELLO FROM B200     ;ourcall:W1AW;"
                               "protocol:DSTAR;sync:voice;yourcall:DL1ABC")
    assert lines[5:] == ["protocol:DSTAR", ""]


CRC_SRC = r"""
#define __host__
#define __device__
#include "crc_par.cuh"
#include <cstdio>
#include <cstdint>
#include <random>
using namespace dh;
template <int KIND, int N> unsigned long long check() {
    constexpr CrcTable<N> t = make_crc_table<KIND, N>();
    std::mt19937_64 rng(KIND * 1000 + N);
    unsigned long long bad = 0;
    for (int trial = 0; trial < 20000; trial++) {
        unsigned char bits[N];
        for (int i = 0; i < N; i++) bits[i] = (trial < 2 ? trial : (rng() >> 17)) & 1;
        // the reference's bit-serial update (crc16.c:3-19 / sacch.cpp:76-90 / facch1.cpp:60-75)
        uint32_t crc = KIND == 0 ? 0u : (KIND == 1 ? 0x3Fu : 0xFFFu);
        for (int i = 0; i < N; i++)
            crc = KIND == 0 ? crc_step_ysf16(crc, bits[i]) : (KIND == 1 ? crc_step_nxdn6(crc, bits[i]) : crc_step_nxdn12(crc, bits[i]));
        if (KIND == 0) crc ^= 0xFFFFu;
        uint32_t par = t.c;
        for (int i = 0; i < N; i++) if (bits[i]) par ^= t.t[i];
        if (par != crc) bad++;
    }
    return bad;
}
int main() {
    unsigned long long bad = check<0, 32>() + check<0, 80>() + check<0, 160>() + check<1, 26>() + check<2, 80>();
    printf("%llu\n", bad);
    return 0;
}
"""


def test_parallel_crc_tables_equal_bit_serial_crc():
    """crc_par.cuh: the compile-time affine tables used by the YSF / NXDN kernels reproduce the bit-serial CRCs of the
    reference for random messages of every length in use (host build of the same header)."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "crc.cpp")
        open(src, "w").write(CRC_SRC)
        exe = os.path.join(d, "crc")
        subprocess.run(["g++", "-std=c++17", "-O1", "-fconstexpr-ops-limit=200000000", "-fconstexpr-loop-limit=1000000",
                        "-I" + os.path.join(ROOT, "digiham_b200", "csrc"), src, "-o", exe], check=True)
        out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert int(out.strip()) == 0


def test_headers_compile_without_a_gpu():
    """include/digiham_b200.h is plain C99; the header-compatible C++ facades compile against the csdr shim."""
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "digiham_b200.h")], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I" + inc,
                    "-I" + os.path.join(ROOT, "oracle", "csdr_shim"), os.path.join(ROOT, "tests", "cpp", "facade_pipe.cpp")],
                   check=True)


def test_synthetic_generators_are_deterministic():
    from digiham_b200 import synth
    a = synth.dmr_symbols(20, seed=3)
    b = synth.dmr_symbols(20, seed=3)
    assert np.array_equal(a, b) and a.max() <= 3
    xa, _ = synth.dmr_channel_bank(3, 5000, seed=1, device="cpu")
    xb, _ = synth.dmr_channel_bank(3, 5000, seed=1, device="cpu")
    assert np.array_equal(xa.numpy(), xb.numpy()) and float(xa.abs().max()) <= 1.0
    # systematic encoders: valid codewords have a zero syndrome in the generated tables' sense
    for code, n in (("hamming_7_4", 4), ("golay_20_8", 8), ("qr_16_7", 7)):
        for d in range(1 << min(n, 6)):
            cw = synth.encode_block(code, d)
            assert cw >> (synth._P[code][0] - synth._P[code][1]) == d


def test_meta_collector_hold_release_semantics():
    """Digiham::MetaCollector of include/meta.hpp behaves like the reference's (include/meta.hpp:50-66,
    src/lib/meta.cpp:58-100): updates without a writer are dropped, hold()/release() coalesce updates into one that
    carries the latest state, the explicit-map overload is not held back, the collector owns (and closes) its writer."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "mc")
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "oracle", "csdr_shim"),
                        os.path.join(ROOT, "tests", "cpp", "meta_collector.cpp"), "-o", exe], check=True)
        out = os.path.join(d, "meta.txt")
        subprocess.run([exe, out], check=True)
        text = open(out).read()
    assert text == ("call:DL1ABC;protocol:PROBE\n" "x:1\n" "call:DL3GHI;protocol:PROBE\n" "call:DL4JKL;protocol:PROBE\n")
