#!/usr/bin/env python3
"""Generate digiham_b200/csrc/tables.inc — the constant tables of the CUDA hot path.

Everything here is derived from first principles, not copied from the reference's C tables:

* RRC taps: the published mkshape design values (wide: 81 taps, narrow: 161 taps; reference
  src/rrc_filter/rrc_filter.cpp:39-115).  Both impulse responses are symmetric, so only the first half
  (incl. centre) is listed, in units of 1e-10, and mirrored.  decimal -> double -> float32 is the same
  conversion chain the reference's `(const float[]){ <double literals> }` goes through.
* Block codes: the parity part P of each systematic generator matrix G = [I | P] as printed in ETSI TS 102
  361-1 annex B.3 (DMR), the YSF spec appendix A (Golay(24,12)) and the POCSAG generator polynomial.  The check
  matrix is H = [P^T | I]; the syndrome has H's row 0 as its MSB (reference hamming_13_9.c:53-72 and siblings).
  The correction LUT is built by enumerating every error pattern of weight <= t in the same order as the
  reference's *_syndrome_generator.c programs and keeping the FIRST pattern seen per syndrome, which is what
  the reference's linear search over `corrections[]` returns.  tests/test_oracle_cpu.py proves LUT equality with
  the compiled reference for every syndrome on the CPU; tests/test_fec_gpu.py drives the device decoders that use
  the tables over all words / error patterns.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "digiham_b200", "csrc", "tables.inc")

# ---------------------------------------------------------------------------------------------------------------
# RRC taps, first half including the centre tap, units of 1e-10
WIDE_HALF = [
    -8938217, -2609230, 5898982, 16095188, 26805019, 35892828, 40255371, 36242975, 20553299, -8516117,
    -49736668, -97942071, -143781385, -174576799, -176417629, -137316693, -50921107, 80011038, 241300735,
    407081846, 542175970, 607228306, 566126484, 394623171, 88613798, -329693214, -809351463, -1273151201,
    -1625361486, -1764143887, -1597076656, -1057455528, -118628528, 1196309860, 2811569136, 4603559944,
    6413467573, 8066010425, 9391765221, 10249723677, 10546584365,
]
NARROW_HALF = [
    -8965127, -6084266, -2629259, 1376901, 5891423, 10840181, 16105739, 21516457, 26838327, 31771176,
    35950725, 38957679, 40334554, 39610403, 36332901, 30106572, 20635228, 7766025, -8467956, -27810092,
    -49751193, -73512625, -98044779, -122043473, -143986008, -162187503, -174876896, -180290597, -176780431,
    -162931143, -137681562, -100442577, -51204456, 9374242, 79903670, 158232514, 241456376, 325968938,
    407558163, 481547523, 542979823, 586838603, 608299644, 603002781, 567332283, 498692532, 395764841,
    258730951, 89449258, -108429006, -329414440, -566213193, -809844704, -1049844817, -1274551627,
    -1471467396, -1627685874, -1730370678, -1767267207, -1727227994, -1600729711, -1380359261, -1061246612,
    -641423317, -122087987, 492236806, 1193667582, 1971049660, 2810174958, 3694123940, 4603722307,
    5518097911, 6415318736, 7273088884, 8069476569, 8783646253, 9396566353, 9891664557, 10255404526,
    10477760738, 10552572221,
]
WIDE_GAIN = "8.337797030e+00"
NARROW_GAIN = "1.667711971e+01"


def taps(half):
    vals = [float("%de-10" % v) for v in half]          # correctly rounded decimal -> double
    full = vals + vals[-2::-1]
    return np.array(full, dtype=np.float64).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# Block codes.  P rows as bit strings, one per data bit (MSB-first data ordering).
CODES = {
    # name: (n, k, P rows, max error weight, enumeration)
    "hamming_7_4": (7, 4, ["101", "111", "110", "011"], 1, "tri"),
    "hamming_13_9": (13, 9, ["1111", "1110", "0111", "1010", "0101", "1011", "1100", "0110", "0011"], 1, "tri"),
    "hamming_15_11": (15, 11, ["1001", "1101", "1111", "1110", "0111", "1010", "0101", "1011", "1100", "0110",
                               "0011"], 1, "tri"),
    "hamming_16_11": (16, 11, ["10011", "11010", "11111", "11100", "01110", "10101", "01011", "10110", "11001",
                               "01101", "00111"], 1, "tri"),
    "qr_16_7": (16, 7, ["001001111", "100011110", "110110111", "111100010", "111001001", "011100101",
                        "001110011"], 2, "full"),
    "golay_20_8": (20, 8, ["001111011010", "110110011001", "011011001101", "001101100111", "110111000110",
                           "101010010111", "100100111110", "100011101011"], 3, "tri"),
    "golay_24_12": (24, 12, ["110001110101", "011000111011", "111101101000", "011110110100", "001111011010",
                             "110110011001", "011011001101", "001101100111", "110111000110", "101010010111",
                             "100100111110", "100011101011"], 3, "tri"),
}

BCH_POLY = 0b11101101001   # x^10 + x^9 + x^8 + x^6 + x^5 + x^3 + 1, POCSAG BCH(31,21)


def h_rows_from_p(n, k, prow):
    r = n - k
    rows = []
    for j in range(r):
        v = 0
        for d in range(k):
            if prow[d][j] == "1":
                v |= 1 << (n - 1 - d)
        v |= 1 << (r - 1 - j)
        rows.append(v)
    return rows


def bch_rows():
    n, r = 31, 10
    cols = []
    for l in range(n):
        # x^l mod g(x)
        v = 1 << l
        for s in range(l, r - 1, -1):
            if v & (1 << s):
                v ^= BCH_POLY << (s - r)
        cols.append(v)
    rows = []
    for j in range(r):
        v = 0
        for l in range(n):
            if cols[l] & (1 << (r - 1 - j)):
                v |= 1 << l
        rows.append(v)
    return rows


def syndrome(rows, w):
    s = 0
    for row in rows:
        s = (s << 1) | (bin(row & w).count("1") & 1)
    return s


def patterns(n, t, mode):
    """Error patterns in the order the reference's generator programs emit them."""
    for i in range(n):
        yield 1 << i
        if t < 2:
            continue
        if mode == "full":
            # quadratic_residue_syndrome_generator.c:27-34: every ordered pair
            for k in range(n):
                if k != i:
                    yield (1 << i) ^ (1 << k)
        else:
            for k in range(i):
                yield (1 << i) ^ (1 << k)
                if t >= 3:
                    for l in range(k):
                        yield (1 << i) ^ (1 << k) ^ (1 << l)


def build_lut(n, rows, t, mode):
    r = len(rows)
    lut = [0] * (1 << r)
    for e in patterns(n, t, mode):
        s = syndrome(rows, e)
        if s != 0 and lut[s] == 0:
            lut[s] = e
    return lut


def c_array(ctype, name, vals, fmt, per_line=8):
    """Emits `#define <NAME>_INIT { ... }` (usable as a __constant__ initialiser) and, unless
    DH_TABLES_NO_HOST_ARRAYS is defined, a static host array initialised from it."""
    macro = name.upper() + "_INIT"
    lines = ["#define %s { \\" % macro]
    for i in range(0, len(vals), per_line):
        lines.append("    " + ", ".join(fmt % v for v in vals[i:i + per_line]) + ", \\")
    lines.append("}")
    lines.append("#define %s_LEN %d" % (name.upper(), len(vals)))
    lines.append("#ifndef DH_TABLES_NO_HOST_ARRAYS")
    lines.append("static const %s %s[%d] = %s;" % (ctype, name, len(vals), macro))
    lines.append("#endif")
    return "\n".join(lines)


def main():
    out = []
    out.append("// GENERATED by tools/gen_tables.py — do not edit.  See that script for provenance.")
    out.append("#pragma once")
    out.append("#include <stdint.h>")
    out.append("")
    for name, half, gain in (("WIDE", WIDE_HALF, WIDE_GAIN), ("NARROW", NARROW_HALF, NARROW_GAIN)):
        t = taps(half)
        out.append("#define DH_RRC_%s_NZEROS %d" % (name, len(t) - 1))
        out.append("#define DH_RRC_%s_GAIN %s" % (name, gain))
        out.append(c_array("uint32_t", "dh_rrc_%s_taps_bits" % name.lower(), t.view(np.uint32).tolist(), "0x%08xu"))
        out.append("")
    codes = dict(CODES)
    specs = []
    for name, (n, k, prow, t, mode) in codes.items():
        specs.append((name, n, h_rows_from_p(n, k, prow), t, mode))
    specs.append(("bch_31_21", 31, bch_rows(), 2, "tri"))
    for name, n, rows, t, mode in specs:
        lut = build_lut(n, rows, t, mode)
        r = len(rows)
        ctype = "uint8_t" if n <= 8 else ("uint16_t" if n <= 16 else "uint32_t")
        out.append("// %s: n=%d, r=%d, %d correctable syndromes" % (name, n, r, sum(1 for v in lut if v)))
        out.append(c_array("uint32_t", "dh_%s_h" % name, rows, "0x%08xu"))
        out.append(c_array(ctype, "dh_%s_lut" % name, lut, "0x%xu", 12))
        out.append("")
    text = "\n".join(out) + "\n"
    with open(OUT, "w") as f:
        f.write(text)
    sys.stderr.write("wrote %s (%d bytes)\n" % (os.path.normpath(OUT), len(text)))


if __name__ == "__main__":
    main()
