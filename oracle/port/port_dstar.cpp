// port_dstar.cpp — CPU restatement of the reference's D-Star decoder incl. its metadata plane.
// TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
//
// Follows Digiham::DStar::{SyncPhase,HeaderPhase,VoicePhase} (reference src/dstar_decoder/dstar_phase.cpp:17-278),
// Header (header.cpp:23-189), Scrambler (scrambler.cpp:6-21), Crc (crc.cpp:6-23) and MetaCollector
// (dstar_meta.cpp:5-130 with the hold/release batching of src/lib/meta.cpp:71-100).
// Two places where the reference's behaviour is undefined are pinned instead: std::stof on a non-numeric NMEA
// field (throws in the reference) and fewer than six GGA fields (out-of-range vector access) skip the sentence.
#include "port.hpp"

#include <cstdlib>
#include <cstring>

namespace port {

namespace {

const uint8_t kHeaderSync[24] = {0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0, 0};
const uint8_t kVoiceSync[24] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0};
const uint8_t kTerminator[48] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0,
                                 1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 0};

void descramble(const uint8_t* in, uint8_t* out, size_t n) {   // scrambler.cpp:10-21, register reset to all ones
    unsigned sr = 0x7F;
    for (size_t i = 0; i < n; i++) {
        const unsigned wb = (sr & 1u) ^ ((sr >> 3) & 1u);
        out[i] = (uint8_t) ((in[i] & 1u) ^ wb);
        sr = ((sr & 0x7Eu) >> 1) | (wb << 6);
    }
}

uint16_t crcOf(const uint8_t* data, size_t len) {   // crc.cpp:6-20
    uint16_t c = 0xFFFF;
    for (size_t k = 0; k < len; k++) {
        for (int i = 0; i < 8; i++) {
            c ^= (data[k] >> i) & 1;
            c = (c & 1) ? (uint16_t) ((c >> 1) ^ 0x8408) : (uint16_t) (c >> 1);
        }
    }
    return (uint16_t) (c ^ 0xFFFF);
}

std::string rtrim(std::string s) {
    s.erase(s.find_last_not_of(' ') + 1);
    return s;
}

}  // namespace

bool dstar_header_crc_ok(const uint8_t h[41]) { return crcOf(h, 39) == (uint16_t) (h[39] | h[40] << 8); }

// Header::parseFromHeader (header.cpp:23-48): descramble, de-interleave, 4-state Viterbi over 330 steps with the
// decoded bits stored LSB first; more than 10 channel errors or a CRC mismatch reject the header.
bool dstar_header_decode(const uint8_t* raw660, uint8_t out[42]) {
    uint8_t d[660], t[660];
    descramble(raw660, d, 660);
    for (int i = 0; i < 12; i++) {
        for (int k = 0; k < 28; k++) t[k * 24 + i] = d[i * 28 + k];
    }
    for (int i = 12; i < 24; i++) {
        for (int k = 0; k < 27; k++) t[k * 24 + i] = d[12 + i * 27 + k];
    }
    auto expected = [](unsigned prev, unsigned b) {   // header.cpp:71-76
        unsigned e = b ? 3u : 0u;
        if (prev & 1) e ^= 3u;
        if (prev & 2) e ^= 2u;
        return e;
    };
    uint16_t metric[4] = {0, 0, 0, 0}, nextMetric[4];
    std::vector<std::vector<uint8_t>> path(4, std::vector<uint8_t>(42, 0)), nextPath(4);
    for (int pos = 0; pos < 330; pos++) {
        const unsigned rx = ((t[2 * pos] & 1u) << 1) | (t[2 * pos + 1] & 1u);
        for (unsigned s = 0; s < 4; s++) {
            const unsigned bit = (s >> 1) & 1;
            unsigned chosen = 0;
            uint16_t best = 0;
            for (unsigned k = 0; k < 2; k++) {
                const unsigned prev = ((s << 1) & 2u) | k;
                const uint16_t m = (uint16_t) (metric[prev] + __builtin_popcount(rx ^ expected(prev, bit)));
                if (k == 0 || m < best) {
                    best = m;
                    chosen = prev;
                }
            }
            nextMetric[s] = best;
            nextPath[s] = path[chosen];
            nextPath[s][pos / 8] |= (uint8_t) (bit << (pos % 8));
        }
        std::memcpy(metric, nextMetric, sizeof(metric));
        path.swap(nextPath);
    }
    unsigned winner = 0;
    for (unsigned s = 1; s < 4; s++) {
        if (metric[s] < metric[winner]) winner = s;
    }
    std::memcpy(out, path[winner].data(), 42);
    if (metric[winner] > 10) return false;
    return dstar_header_crc_ok(out);
}

namespace {

struct DStar {
    Decoded* out = nullptr;
    // MetaCollector (dstar_meta.hpp:26-33)
    std::string sync, message, departure, destination, ourCall, yourCall, dprs;
    bool located = false;
    float lat = 0, lon = 0;
    int held = 0;
    bool dirty = false;
    // VoicePhase (dstar_phase.hpp:62-73)
    int frameCount = 0, syncCount = 0;
    uint8_t collected[6] = {0};
    uint8_t msg[20] = {0};
    unsigned msgBlocks = 0;
    uint8_t hdr[41] = {0};
    unsigned hdrCount = 0;
    std::string simpleData;

    void send() {
        if (held) {
            dirty = true;
            return;
        }
        std::map<std::string, std::string> kv;
        kv["protocol"] = "DSTAR";
        if (!sync.empty()) kv["sync"] = sync;
        if (!departure.empty()) kv["departure"] = departure;
        if (!destination.empty()) kv["destination"] = destination;
        if (!ourCall.empty()) kv["ourcall"] = ourCall;
        if (!yourCall.empty()) kv["yourcall"] = yourCall;
        if (!message.empty()) kv["message"] = message;
        if (!dprs.empty()) kv["dprs"] = dprs;
        if (located) {
            kv["lat"] = std::to_string(lat);
            kv["lon"] = std::to_string(lon);
        }
        out->meta += serialize(kv);
    }
    void set(std::string& f, const std::string& v) {
        if (f == v) return;
        f = v;
        send();
    }
    void setGps(bool valid, float la, float lo) {   // dstar_meta.cpp:62-72
        if (!valid && !located) return;
        if (valid && located && lat == la && lon == lo) return;
        located = valid;
        lat = la;
        lon = lo;
        send();
    }
    void release() {
        if (--held == 0) {
            if (dirty) send();
            dirty = false;
        }
    }
    void resetMeta() {   // dstar_meta.cpp:74-85
        held++;
        set(sync, "");
        set(message, "");
        set(departure, "");
        set(destination, "");
        set(ourCall, "");
        set(yourCall, "");
        set(dprs, "");
        setGps(false, 0, 0);
        release();
    }
    void setFromHeader(const uint8_t* h) {   // dstar_meta.cpp:15-27, header.cpp:150-178
        held++;
        set(sync, (h[0] >> 7) & 1 ? "data" : "voice");
        set(departure, rtrim(latin1_to_utf8(h + 11, 8)));
        set(destination, rtrim(latin1_to_utf8(h + 3, 8)));
        std::string own = rtrim(latin1_to_utf8(h + 27, 8));
        const std::string suffix = rtrim(latin1_to_utf8(h + 35, 4));
        if (suffix != "") own += "/" + suffix;
        set(ourCall, own);
        set(yourCall, rtrim(latin1_to_utf8(h + 19, 8)));
        release();
    }

    void startVoice(int frames, int syncs) {
        std::memset(collected, 0, 6);
        resetFrames();
        simpleData.clear();
        frameCount = frames;
        syncCount = syncs;
    }
    void resetFrames() {   // dstar_phase.cpp:155-161
        frameCount = 0;
        std::memset(msg, 0, 20);
        msgBlocks = 0;
        std::memset(hdr, 0, 41);
        hdrCount = 0;
    }
    void collectDataFrame(const uint8_t* d3) {   // dstar_phase.cpp:163-211
        std::memcpy(collected + (frameCount % 2) * 3, d3, 3);
        if (frameCount % 2 == 0) return;
        const unsigned n = collected[0] & 0x0F;
        switch (collected[0] >> 4) {
            case 4:
                if (n > 3) break;
                std::memcpy(msg + n * 5, collected + 1, 5);
                msgBlocks |= 1u << n;
                break;
            case 5:
                if (n > 5 || hdrCount + n > 41) break;
                std::memcpy(hdr + hdrCount, collected + 1, n);
                hdrCount += n;
                break;
            case 3:
                if (n > 5) break;
                simpleData += std::string((const char*) collected + 1, n);
                break;
            default: break;
        }
    }
    void parseNmea(const std::string& input) {   // dstar_phase.cpp:245-278
        const size_t star = input.find_last_of('*');
        if (star == std::string::npos || star + 2 > input.length()) return;
        const std::string body = input.substr(1, star - 1);
        if (body.length() < 2) return;   // reference: substr(2, 3) would throw
        const std::string sentence = body.substr(2, 3);
        uint8_t checksum = 0;
        for (char ch : body) checksum ^= (uint8_t) ch;
        const unsigned toCheck = (unsigned) std::strtoul(input.substr(star + 1, 2).c_str(), nullptr, 16);
        if (checksum != toCheck) return;
        std::vector<std::string> fields;
        size_t from = 0;
        while (from <= body.length()) {   // getline(',') semantics: no trailing empty field
            const size_t comma = body.find(',', from);
            if (comma == std::string::npos) {
                if (from < body.length()) fields.push_back(body.substr(from));
                break;
            }
            fields.push_back(body.substr(from, comma - from));
            from = comma + 1;
        }
        if (sentence != "GGA" || fields.size() < 6) return;
        char* end = nullptr;
        const float latC = std::strtof(fields[2].c_str(), &end);
        if (end == fields[2].c_str()) return;
        const float lonC = std::strtof(fields[4].c_str(), &end);
        if (end == fields[4].c_str()) return;
        float la = (float) ((int) latC / 100);
        la += (latC - la * 100) / 60;
        if (fields[3] == "S") la *= -1;
        float lo = (float) ((int) lonC / 100);
        lo += (lonC - lo * 100) / 60;
        if (fields[5] == "W") lo *= -1;
        setGps(true, la, lo);
    }
    void parseFrameData() {   // dstar_phase.cpp:213-243
        if (msgBlocks == 0x0F) set(message, latin1_to_utf8(msg, 20));
        if (hdrCount == 41 && dstar_header_crc_ok(hdr)) setFromHeader(hdr);
        size_t pos;
        while ((pos = simpleData.find('\r')) != std::string::npos) {
            const std::string s = simpleData.substr(0, pos + 1);
            if (s.length() >= 10 && s.substr(0, 5) == "$$CRC" && s.at(9) == ',') {
                const unsigned check = (unsigned) std::strtoul(s.substr(5, 4).c_str(), nullptr, 16);
                if (crcOf((const uint8_t*) s.data() + 10, s.length() - 10) == (uint16_t) check)
                    set(dprs, s.substr(10, s.length() - 11));
            } else if (s.length() > 5 && s.at(0) == '$') {
                parseNmea(s);
            }
            simpleData = simpleData.substr(pos + 1 + (simpleData.length() > pos + 1 && simpleData.at(pos + 1) == '\n'));
        }
    }
};

}  // namespace

void decode_dstar(const uint8_t* sym, size_t n, Decoded& out) {
    DStar d;
    d.out = &out;
    enum { kSync, kHeader, kVoice } phase = kSync;
    size_t pos = 0;
    for (;;) {
        const uint8_t* p = sym + pos;
        if (phase == kSync) {   // dstar_phase.cpp:17-35
            if (n - pos <= 24) break;
            if (hamming_distance(p, kHeaderSync, 24) <= 2) {
                pos += 24;
                phase = kHeader;
            } else if (hamming_distance(p, kVoiceSync, 24) <= 1) {
                pos += 24;
                d.startVoice(0, 0);
                phase = kVoice;
            } else {
                pos++;
            }
        } else if (phase == kHeader) {   // dstar_phase.cpp:37-59
            if (n - pos <= 660) break;
            uint8_t h[42];
            if (!dstar_header_decode(p, h)) {
                pos += 1;
                phase = kSync;
                continue;
            }
            pos += 660;
            if (!((h[0] >> 7) & 1)) {
                d.setFromHeader(h);
                d.startVoice(21, 1);
                phase = kVoice;
            } else {
                phase = kSync;
            }
        } else {   // VoicePhase::process, dstar_phase.cpp:78-149
            if (n - pos <= 120) break;
            if (d.syncCount >= 1) {
                uint8_t v[9] = {0};
                for (int i = 0; i < 72; i++) v[i / 8] |= (uint8_t) ((p[i] & 1) << (i % 8));
                out.bytes.insert(out.bytes.end(), v, v + 9);
            }
            const uint8_t* data = p + 72;
            pos += 96;
            if (hamming_distance(data, kTerminator, 48) <= 1 || hamming_distance(data, kTerminator + 24, 24) <= 1) {
                pos += 24;
                d.resetMeta();
                phase = kSync;
                continue;
            }
            if (d.frameCount >= 20) {
                if (hamming_distance(data, kVoiceSync, 24) > 1) {
                    if (--d.syncCount < 0) {
                        d.resetMeta();
                        phase = kSync;
                        continue;
                    }
                } else {
                    if (++d.syncCount > 3) d.syncCount = 3;
                    if (d.syncCount > 1) d.set(d.sync, "voice");
                }
                d.parseFrameData();
                d.resetFrames();
            } else {
                uint8_t bits[24], bytes[3] = {0, 0, 0};
                descramble(data, bits, 24);
                for (int i = 0; i < 24; i++) bytes[i / 8] |= (uint8_t) (bits[i] << (i % 8));
                d.collectDataFrame(bytes);
                d.frameCount++;
            }
        }
    }
}

int dstar_header_probe(const uint8_t* raw660, char* text, size_t cap) {
    uint8_t h[42];
    if (!dstar_header_decode(raw660, h)) return -1;
    auto field = [&](int off, int len) { return rtrim(latin1_to_utf8(h + off, (size_t) len)); };
    std::string own = field(27, 8);
    if (field(35, 4) != "") own += "/" + field(35, 4);
    const std::string s = "DST RPT: \"" + field(3, 8) + "\" DPT RPT: \"" + field(11, 8) + "\" COMPANION: \"" +
                          field(19, 8) + "\" CALLSIGN: \"" + own + "\" ";
    if (text && cap) {
        const size_t m = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(text, s.data(), m);
        text[m] = 0;
    }
    return (h[0] >> 7) & 1;
}

}  // namespace port
