// port_nxdn.cpp — CPU restatement of the reference's NXDN decoder incl. its metadata plane.
// TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
//
// Follows Digiham::Nxdn::{SyncPhase,FramedPhase} (reference src/nxdn_decoder/nxdn_phase.cpp:19-171), Scrambler
// (scrambler.cpp:8-25), Lich (lich.cpp:5-49), Sacch / SacchSuperframeCollector / SacchSuperframe
// (sacch.cpp:24-154), Facch1 (facch1.cpp:8-75), Trellis (trellis.cpp:10-101) and MetaCollector
// (nxdn_meta.cpp:6-76 with the hold/release batching of src/lib/meta.cpp:71-100).
#include "port.hpp"

#include <cstring>

namespace port {

// reference src/nxdn_decoder/trellis.cpp:29-101: rate 1/2, K = 5, hard decision, register exchange, 16-bit
// metrics, and the start-state prior: while `blocked` (0b1111 shifted left once per step) intersects a state's
// index, only its predecessor k = 0 is evaluated.  in: nbits/2 dibits, 4 per byte, MSB first.
unsigned nxdn_viterbi(const uint8_t* in, unsigned nbits, uint8_t* out) {
    const unsigned bytes = (nbits + 15) / 16;
    auto expected = [](unsigned prev, unsigned b) {
        const unsigned d1 = (prev >> 3) & 1, d2 = (prev >> 2) & 1, d3 = (prev >> 1) & 1, d4 = prev & 1;
        return ((b ^ d3 ^ d4) << 1) | (b ^ d1 ^ d2 ^ d4);
    };
    std::vector<uint16_t> metric(16, 0), nextMetric(16);
    std::vector<std::vector<uint8_t>> path(16, std::vector<uint8_t>(bytes, 0)), nextPath(16);
    unsigned blocked = 15;
    for (unsigned pos = 0; pos < nbits / 2; pos++) {
        const unsigned rx = (in[pos / 4] >> (2 * (3 - pos % 4))) & 3u;
        for (unsigned s = 0; s < 16; s++) {
            const unsigned bit = (s >> 3) & 1;
            const unsigned candidates = (s & blocked) == 0 ? 2 : 1;
            unsigned chosen = 0;
            uint16_t best = 0;
            for (unsigned k = 0; k < candidates; k++) {
                const unsigned prev = ((s << 1) & 14u) | k;
                const uint16_t m = (uint16_t) (metric[prev] + __builtin_popcount(rx ^ expected(prev, bit)));
                if (k == 0 || m < best) {
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
        blocked = (blocked << 1) & 15u;
    }
    unsigned winner = 0;
    for (unsigned s = 1; s < 16; s++) {
        if (metric[s] < metric[winner]) winner = s;
    }
    std::memcpy(out, path[winner].data(), bytes);
    return metric[winner];
}

namespace {

const uint8_t kFsw[10] = {3, 0, 3, 1, 3, 3, 1, 1, 2, 1};   // nxdn_phase.cpp:17

unsigned bitOf(const uint8_t* dibits, unsigned pos) { return (dibits[pos / 2] >> (1 - pos % 2)) & 1u; }
unsigned msbBit(const uint8_t* bytes, unsigned pos) { return (bytes[pos / 8] >> (7 - pos % 8)) & 1u; }

// generic "de-interleave rows x cols, re-insert punctured zeros, Viterbi": Sacch::parse / Facch1::parse front half
template <typename Punctured>
void channelDecode(const uint8_t* dibits, unsigned rows, unsigned cols, unsigned codedBits, Punctured punctured,
                   uint8_t* decoded) {
    std::vector<uint8_t> stream(rows * cols);
    for (unsigned i = 0; i < rows; i++) {
        for (unsigned k = 0; k < cols; k++) stream[k * rows + i] = (uint8_t) bitOf(dibits, i * cols + k);
    }
    std::vector<uint8_t> packed((codedBits + 7) / 8, 0);
    unsigned next = 0;
    for (unsigned i = 0; i < codedBits; i++) {
        unsigned x = 0;
        if (!punctured((int) i)) x = stream[next++];
        packed[i / 8] |= (uint8_t) (x << (7 - i % 8));
    }
    nxdn_viterbi(packed.data(), codedBits, decoded);
}

unsigned serialCrc(const uint8_t* bytes, unsigned nbits, unsigned width, unsigned poly, unsigned init) {
    unsigned crc = init;
    const unsigned mask = (1u << width) - 1;
    for (unsigned i = 0; i < nbits; i++) {
        const unsigned cb = ((crc >> (width - 1)) & 1u) ^ msbBit(bytes, i);
        if (cb) crc ^= poly;
        crc = ((crc << 1) & (mask & ~1u)) | cb;
    }
    return crc;
}

struct Nxdn {
    Decoded* out = nullptr;
    // MetaCollector (nxdn_meta.hpp:20-23)
    std::string sync, type;
    unsigned source = 0, destination = 0;
    int held = 0;
    bool dirty = false;
    // FramedPhase (nxdn_phase.hpp:37-40)
    bool framed = false;
    int syncCount = 0;
    int lich = -1;
    bool have[4] = {false, false, false, false};
    uint8_t fragment[4][5];

    void send() {
        if (held) {
            dirty = true;
            return;
        }
        std::map<std::string, std::string> kv;
        kv["protocol"] = "NXDN";
        if (!sync.empty()) kv["sync"] = sync;
        if (!type.empty()) kv["type"] = type;
        if (source != 0) kv["source"] = std::to_string(source);
        if (destination != 0) kv["destination"] = std::to_string(destination);
        out->meta += serialize(kv);
    }
    void setStr(std::string& f, const std::string& v) {
        if (f == v) return;
        f = v;
        send();
    }
    void setNum(unsigned& f, unsigned v) {
        if (f == v) return;
        f = v;
        send();
    }
    void resetMeta() {   // nxdn_meta.cpp:68-75
        held++;
        setStr(sync, "");
        setStr(type, "");
        setNum(source, 0);
        setNum(destination, 0);
        if (--held == 0) {
            if (dirty) send();
            dirty = false;
        }
    }

    static bool isSync(const uint8_t* p) { return hamming_distance(p, kFsw, 10) <= 2; }

    void enterFramed() {
        framed = true;
        syncCount = 0;
        lich = -1;
        for (bool& h : have) h = false;
    }

    void sacch(const uint8_t* dibits) {   // sacch.cpp:24-43 + collector :96-134 + nxdn_phase.cpp:103-115
        uint8_t d[5];
        channelDecode(dibits, 12, 5, 72, [](int i) { return (i + 1) % 6 == 0; }, d);
        if ((d[3] & 0x3Fu) != serialCrc(d, 26, 6, 0x13, 0x3F)) return;
        const unsigned index = (d[0] >> 6) ^ 3u;
        if (index > 0 && !have[index - 1]) return;
        std::memcpy(fragment[index], d, 5);
        have[index] = true;
        if (!(have[0] && have[1] && have[2] && have[3])) return;
        uint8_t sf[9] = {0};
        for (unsigned i = 0; i < 4; i++) {
            for (unsigned k = 0; k < 18; k++) {
                const unsigned o = i * 18 + k;
                sf[o / 8] |= (uint8_t) (msbBit(fragment[i] + 1, k) << (7 - o % 8));
            }
        }
        if ((sf[0] & 0x3Fu) == 0x01) {   // VCALL -> MetaCollector::setFromSacch (nxdn_meta.cpp:54-66)
            const unsigned callType = sf[2] >> 5;
            setStr(type, callType == 1 ? "conference" : (callType == 4 ? "individual" : ""));
            setNum(source, (unsigned) sf[3] << 8 | sf[4]);
            setNum(destination, (unsigned) sf[5] << 8 | sf[6]);
        }
        for (bool& h : have) h = false;
    }

    int facch1(const uint8_t* dibits) {   // facch1.cpp:8-26
        uint8_t d[12];
        channelDecode(dibits, 16, 9, 192, [](int i) { return (i - 1) % 4 == 0; }, d);
        const unsigned check = (unsigned) d[10] << 4 | d[11] >> 4;
        if (check != serialCrc(d, 80, 12, 0x407, 0xFFF)) return -1;
        return d[0] & 0x3F;
    }

    // FramedPhase::process; returns symbols consumed
    size_t frame(const uint8_t* p) {
        if (isSync(p)) {
            if (++syncCount > 6) syncCount = 6;
        } else if (--syncCount < 0) {
            resetMeta();
            framed = false;
            return 0;
        }
        uint8_t body[182];
        unsigned sr = 0xE4;
        for (int i = 0; i < 182; i++) {   // scrambler.cpp:12-25
            const unsigned wb = sr & 1u;
            body[i] = (uint8_t) ((p[10 + i] & 3u) ^ (wb << 1));
            const unsigned fb = ((sr >> 4) & 1u) ^ wb;
            sr = ((sr & 0x1FEu) >> 1) | (fb << 8);
        }
        unsigned bits[8];
        for (int i = 0; i < 8; i++) bits[i] = (body[i] >> 1) & 1u;
        if (bits[7] == (bits[0] ^ bits[1] ^ bits[2] ^ bits[3])) {
            lich = 0;
            for (int i = 0; i < 7; i++) lich |= (int) (bits[i] << (6 - i));
        }
        if (lich < 0) return 192;
        const int rf = (lich >> 5) & 3, functional = (lich >> 3) & 3, option = (lich >> 1) & 3;
        if (rf == 0 || functional == 1) return 192;
        if (functional == 2) sacch(body + 8);
        for (int i = 0; i < 2; i++) {
            const uint8_t* half = body + 38 + 72 * i;
            if ((option >> (1 - i)) & 1) {
                if (syncCount >= 1) {
                    setStr(sync, "voice");
                    uint8_t v[18] = {0};
                    for (int k = 0; k < 72; k++) v[k / 4] |= (uint8_t) ((half[k] & 3) << (6 - (k % 4) * 2));
                    out->bytes.insert(out->bytes.end(), v, v + 18);
                }
            } else if (facch1(half) == 0x08) {   // TX_RELEASE: this block stays unconsumed (nxdn_phase.cpp:149-152)
                resetMeta();
                framed = false;
                return (size_t) (48 + 72 * i);
            }
        }
        return 192;
    }
};

}  // namespace

void decode_nxdn(const uint8_t* sym, size_t n, Decoded& out) {
    Nxdn d;
    d.out = &out;
    size_t pos = 0;
    for (;;) {
        if (!d.framed) {
            if (n - pos <= 10) break;
            if (Nxdn::isSync(sym + pos)) d.enterFramed();
            else pos++;
        } else {
            if (n - pos <= 192) break;
            pos += d.frame(sym + pos);
        }
    }
}

int nxdn_sacch_probe(const uint8_t* dibits30, uint8_t out5[5]) {
    channelDecode(dibits30, 12, 5, 72, [](int i) { return (i + 1) % 6 == 0; }, out5);
    return (out5[3] & 0x3Fu) == serialCrc(out5, 26, 6, 0x13, 0x3F);
}

int nxdn_facch1_probe(const uint8_t* dibits72) {
    Nxdn d;
    return d.facch1(dibits72);
}

}  // namespace port
