// port_fec.cpp — CPU restatement of the reference's block codes, BPTC, Viterbi, CRC and whitening.
// TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
//
// Block codes (reference src/dmr_decoder/hamming_*.c, golay_20_8.c, quadratic_residue.c,
// src/ysf_decoder/golay_24_12.c, src/pocsag_decoder/bch_31_21.c): every decoder computes the syndrome
// H * word (row 0 = MSB) and looks it up in a list of {syndrome, error pattern} built from all error patterns up to
// the code's correction capability; the first list entry with that syndrome wins.  Here the lists are rebuilt at
// start-up from the systematic generator matrices G = [I | P] of the standards (H = [P^T | I]) by the same
// enumeration the reference's *_syndrome_generator.c programs use.
#include "port.hpp"

#include <cstring>
#include <mutex>

namespace port {

namespace {

struct Code {
    int n = 0, r = 0;
    std::vector<uint32_t> rows;                          // H, row 0 first
    std::vector<std::pair<uint32_t, uint32_t>> fixes;    // {syndrome, pattern} in generator order
};

Code g_codes[kNumCodes];
std::once_flag g_once;

uint32_t syn(const Code& c, uint32_t w) {
    uint32_t s = 0;
    for (uint32_t row : c.rows) s = (s << 1) | (uint32_t) (__builtin_popcount(row & w) & 1);
    return s;
}

void fromParity(Code& c, int n, int k, std::initializer_list<const char*> p) {
    c.n = n;
    c.r = n - k;
    c.rows.assign(c.r, 0);
    int d = 0;
    for (const char* row : p) {
        for (int j = 0; j < c.r; j++) {
            if (row[j] == '1') c.rows[j] |= 1u << (n - 1 - d);
        }
        d++;
    }
    for (int j = 0; j < c.r; j++) c.rows[j] |= 1u << (c.r - 1 - j);
}

// enumeration order of the reference's syndrome generators: i, then (i,k<i), then (i,k,l<k); the QR generator
// walks all ordered pairs instead (quadratic_residue_syndrome_generator.c:27-34)
void enumerate(Code& c, int weight, bool orderedPairs) {
    auto add = [&](uint32_t e) { c.fixes.emplace_back(syn(c, e), e); };
    for (int i = 0; i < c.n; i++) {
        add(1u << i);
        if (weight < 2) continue;
        if (orderedPairs) {
            for (int k = 0; k < c.n; k++) {
                if (k != i) add((1u << i) ^ (1u << k));
            }
            continue;
        }
        for (int k = 0; k < i; k++) {
            add((1u << i) ^ (1u << k));
            if (weight < 3) continue;
            for (int l = 0; l < k; l++) add((1u << i) ^ (1u << k) ^ (1u << l));
        }
    }
}

void build() {
    // ETSI TS 102 361-1 B.3.5, B.3.4 (x3), B.3.2, B.3.1; YSF spec appendix A; POCSAG generator polynomial
    fromParity(g_codes[H7_4], 7, 4, {"101", "111", "110", "011"});
    fromParity(g_codes[H13_9], 13, 9, {"1111", "1110", "0111", "1010", "0101", "1011", "1100", "0110", "0011"});
    fromParity(g_codes[H15_11], 15, 11,
               {"1001", "1101", "1111", "1110", "0111", "1010", "0101", "1011", "1100", "0110", "0011"});
    fromParity(g_codes[H16_11], 16, 11,
               {"10011", "11010", "11111", "11100", "01110", "10101", "01011", "10110", "11001", "01101", "00111"});
    fromParity(g_codes[QR16_7], 16, 7,
               {"001001111", "100011110", "110110111", "111100010", "111001001", "011100101", "001110011"});
    fromParity(g_codes[GOLAY20_8], 20, 8,
               {"001111011010", "110110011001", "011011001101", "001101100111", "110111000110", "101010010111",
                "100100111110", "100011101011"});
    fromParity(g_codes[GOLAY24_12], 24, 12,
               {"110001110101", "011000111011", "111101101000", "011110110100", "001111011010", "110110011001",
                "011011001101", "001101100111", "110111000110", "101010010111", "100100111110", "100011101011"});
    {
        // BCH(31,21): H column for bit l is x^l mod g(x), g = x^10+x^9+x^8+x^6+x^5+x^3+1
        Code& c = g_codes[BCH31_21];
        c.n = 31;
        c.r = 10;
        c.rows.assign(10, 0);
        const uint32_t poly = 0x769;
        for (int l = 0; l < 31; l++) {
            uint32_t v = 1u << l;
            for (int s = l; s >= 10; s--) {
                if (v & (1u << s)) v ^= poly << (s - 10);
            }
            for (int j = 0; j < 10; j++) {
                if (v & (1u << (9 - j))) c.rows[j] |= 1u << l;
            }
        }
    }
    enumerate(g_codes[H7_4], 1, false);
    enumerate(g_codes[H13_9], 1, false);
    enumerate(g_codes[H15_11], 1, false);
    enumerate(g_codes[H16_11], 1, false);
    enumerate(g_codes[QR16_7], 2, true);
    enumerate(g_codes[GOLAY20_8], 3, false);
    enumerate(g_codes[GOLAY24_12], 3, false);
    enumerate(g_codes[BCH31_21], 2, false);
}

const Code& code(int id) {
    std::call_once(g_once, build);
    return g_codes[id];
}

}  // namespace

uint32_t syndrome(int id, uint32_t word) { return syn(code(id), word); }

bool correct(int id, uint32_t& word) {
    const Code& c = code(id);
    const uint32_t s = syn(c, word);
    if (s == 0) return true;
    for (const auto& f : c.fixes) {          // linear search, first match wins (hamming_13_9.c:74-84)
        if (f.first == s) {
            word ^= f.second;
            return true;
        }
    }
    return false;
}

// reference src/dmr_decoder/bptc_196_96.c:5-59
bool bptc_196_96(const uint8_t payload[25], uint8_t out[12]) {
    auto bitOf = [&](const uint8_t* p, int i) { return (p[i / 8] >> (7 - i % 8)) & 1; };
    uint8_t plain[196];
    for (int i = 0; i < 196; i++) plain[i] = (uint8_t) bitOf(payload, (i * 181) % 196);   // de-interleave (:12-15)
    uint32_t column[15];
    bool ok = true;
    for (int c = 0; c < 15; c++) {
        uint32_t w = 0;
        for (int k = 0; k < 13; k++) w |= (uint32_t) plain[k * 15 + c + 1] << (12 - k);   // skip R(3) (:23-26)
        ok &= correct(H13_9, w);
        column[c] = w;
    }
    if (!ok) return false;
    uint32_t row[9];
    for (int r = 0; r < 9; r++) {
        uint32_t w = 0;
        for (int k = 0; k < 15; k++) w |= ((column[k] >> (12 - r)) & 1u) << (14 - k);
        ok &= correct(H15_11, w);
        row[r] = w;
    }
    if (!ok) return false;
    // info bits: row 0 columns 3..10, rows 1..8 columns 0..10, MSB first (:45-56)
    int outBit = 0;
    std::memset(out, 0, 12);
    for (int r = 0; r < 9; r++) {
        for (int c = (r == 0 ? 3 : 0); c < 11; c++) {
            const uint32_t b = (row[r] >> (14 - c)) & 1u;
            out[outBit / 8] |= (uint8_t) (b << (7 - outBit % 8));
            outBit++;
        }
    }
    return true;
}

// reference src/ysf_decoder/trellis.c:8-109: rate 1/2, K = 5, hard decision, register exchange, uint8 metrics
unsigned viterbi(const uint8_t* in, unsigned steps, uint8_t* out) {
    const unsigned bytes = (steps + 7) / 8;
    // expected dibit for the transition out of `prev` with input bit b: g1 = b^D3^D4, g2 = b^D1^D2^D4 with the
    // newest bit of the 4-bit state at the MSB (trellis.c:8-25)
    auto expected = [](unsigned prev, unsigned b) {
        const unsigned d1 = (prev >> 3) & 1, d2 = (prev >> 2) & 1, d3 = (prev >> 1) & 1, d4 = prev & 1;
        return ((b ^ d3 ^ d4) << 1) | (b ^ d1 ^ d2 ^ d4);
    };
    std::vector<uint8_t> metric(16, 0), nextMetric(16);
    std::vector<std::vector<uint8_t>> path(16, std::vector<uint8_t>(bytes, 0)), nextPath(16);
    for (unsigned pos = 0; pos < steps; pos++) {
        const unsigned rx = (in[pos / 4] >> (2 * (3 - pos % 4))) & 3u;
        for (unsigned s = 0; s < 16; s++) {
            const unsigned bit = (s >> 3) & 1;
            unsigned chosen = 0;
            uint8_t best = 0;
            for (unsigned k = 0; k < 2; k++) {
                const unsigned prev = ((s << 1) & 14u) | k;
                const uint8_t m = (uint8_t) (metric[prev] + __builtin_popcount(rx ^ expected(prev, bit)));
                if (k == 0 || m < best) {         // ties keep k = 0 (:71)
                    best = m;
                    chosen = prev;
                }
            }
            nextMetric[s] = best;
            nextPath[s] = path[chosen];
            nextPath[s][pos / 8] |= (uint8_t) (bit << (7 - pos % 8));
        }
        metric.swap(nextMetric);
        path.swap(nextPath);
    }
    unsigned winner = 0;
    for (unsigned s = 1; s < 16; s++) {
        if (metric[s] < metric[winner]) winner = s;   // lowest index among equals (:95-98)
    }
    std::memcpy(out, path[winner].data(), bytes);
    return metric[winner];
}

// reference src/ysf_decoder/crc16.c:3-19
uint16_t crc16(const uint8_t* data, int count) {
    uint16_t crc = 0;
    for (int k = 0; k < count; k++) {
        for (int i = 7; i >= 0; i--) {
            const unsigned fb = ((data[k] >> i) & 1u) ^ ((crc >> 15) & 1u);
            crc = (uint16_t) (crc << 1);
            if (fb) crc ^= 0x1021;
        }
    }
    return (uint16_t) (crc ^ 0xFFFF);
}

// reference src/ysf_decoder/whitening.c:6-22
void dewhiten(const uint8_t* in, uint8_t* out, unsigned nbits) {
    unsigned reg = 0x1C9;
    std::memset(out, 0, (nbits + 7) / 8);
    for (unsigned i = 0; i < nbits; i++) {
        const unsigned w = reg & 1u;
        const unsigned bit = ((in[i / 8] >> (7 - i % 8)) & 1u) ^ w;
        out[i / 8] |= (uint8_t) (bit << (7 - i % 8));
        const unsigned fb = ((reg >> 4) & 1u) ^ w;
        reg = (reg >> 1) | (fb << 8);
    }
}

// reference src/lib/hamming_distance.c:3-10
unsigned hamming_distance(const uint8_t* a, const uint8_t* b, size_t n) {
    unsigned d = 0;
    for (size_t i = 0; i < n; i++) d += (unsigned) __builtin_popcount((unsigned) (a[i] ^ b[i]));
    return d;
}

// reference src/lib/meta.cpp:8-17
std::string serialize(const std::map<std::string, std::string>& kv) {
    std::string s;
    for (auto it = kv.begin(); it != kv.end(); ++it) {
        if (it != kv.begin()) s += ";";
        s += it->first + ":" + it->second;
    }
    return s + "\n";
}

// stands in for src/lib/charset.cpp (ICU): iso-8859-1 -> utf-8, result ends at the first NUL
std::string latin1_to_utf8(const unsigned char* p, size_t n) {
    std::string r;
    for (size_t i = 0; i < n && p[i]; i++) {
        if (p[i] < 0x80) {
            r += (char) p[i];
        } else {
            r += (char) (0xC0 | (p[i] >> 6));
            r += (char) (0x80 | (p[i] & 0x3F));
        }
    }
    return r;
}

}  // namespace port
