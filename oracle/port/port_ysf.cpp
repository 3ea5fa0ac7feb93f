// port_ysf.cpp — CPU restatement of the reference's YSF decoder incl. its metadata plane.
// TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
//
// Follows Digiham::Ysf::{SyncPhase,FramePhase} (reference src/ysf_decoder/ysf_phase.cpp:16-361), Fich
// (fich.cpp:12-66), DataCollector/DataFrame (data.cpp:15-88), Gps (gps.cpp:7-105) and MetaCollector
// (ysf_meta.cpp:7-105 with the hold/release batching of src/lib/meta.cpp:71-100).
#include "port.hpp"

#include <cstring>

namespace port {

namespace {

const uint8_t kSync[20] = {3, 1, 1, 0, 1, 3, 0, 1, 3, 0, 2, 1, 1, 2, 0, 3, 1, 0, 3, 1};   // D471C9634D

void packDibits(const uint8_t* dibits, int count, uint8_t* out) {
    std::memset(out, 0, (size_t) (count + 3) / 4);
    for (int i = 0; i < count; i++) out[i / 4] |= (uint8_t) ((dibits[i] & 3) << (6 - 2 * (i % 4)));
}

struct Ysf {
    Decoded* out = nullptr;
    // metadata collector
    std::string mode, target, source, up, down;
    bool located = false;
    float lat = 0, lon = 0;
    int held = 0;
    bool dirty = false;
    // FramePhase members (ysf_phase.hpp:53-56)
    bool framing = false;
    int syncCount = 0;
    bool haveFich = false;
    uint32_t fich = 0;
    bool expectSubFrame = false;
    uint8_t dt[20] = {0};
    unsigned dtNext = 0;

    void send() {
        if (held) {
            dirty = true;
            return;
        }
        std::map<std::string, std::string> kv;
        kv["protocol"] = "YSF";
        if (!mode.empty()) kv["mode"] = mode;
        if (!target.empty()) kv["target"] = target;
        if (!source.empty()) kv["source"] = source;
        if (!up.empty()) kv["up"] = up;
        if (!down.empty()) kv["down"] = down;
        if (located) {
            kv["lat"] = std::to_string(lat);
            kv["lon"] = std::to_string(lon);
        }
        out->meta += serialize(kv);
    }
    void set(std::string& field, const std::string& v) {
        if (field == v) return;
        field = v;
        send();
    }
    void setPosition(bool valid, float la, float lo) {
        if (!valid && !located) return;
        if (valid && located && la == lat && lo == lon) return;
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
    void resetMeta() {
        held++;
        set(mode, "");
        set(target, "");
        set(source, "");
        set(up, "");
        set(down, "");
        setPosition(false, 0, 0);
        release();
    }

    // FramePhase::treatYsfString (ysf_phase.cpp:351-361)
    static std::string callsign(const uint8_t* p) {
        size_t len = 10;
        for (char stop : {'\n', ' '}) {
            const void* hit = std::memchr(p, stop, len);
            if (hit) len = (size_t) ((const uint8_t*) hit - p);
        }
        return latin1_to_utf8(p, len);
    }

    // Fich::parse (fich.cpp:12-52)
    bool parseFich(const uint8_t* d, uint32_t& value) {
        uint8_t dibits[100], packed[25], decoded[13];
        for (int i = 0; i < 100; i++) dibits[i] = d[(i * 20) % 100 + (i * 20) / 100];
        packDibits(dibits, 100, packed);
        viterbi(packed, 100, decoded);
        uint32_t g[4];
        bool ok = true;
        for (int i = 0; i < 4; i++) {
            g[i] = (uint32_t) decoded[3 * i] << 16 | (uint32_t) decoded[3 * i + 1] << 8 | decoded[3 * i + 2];
            ok &= correct(GOLAY24_12, g[i]);
        }
        if (!ok) return false;
        const uint32_t data = (g[0] >> 12) << 20 | (g[1] >> 12) << 8 | (g[2] >> 16);
        const uint16_t check = (uint16_t) ((g[2] & 0xF000u) | (g[3] >> 12));
        const uint8_t be[4] = {(uint8_t) (data >> 24), (uint8_t) (data >> 16), (uint8_t) (data >> 8), (uint8_t) data};
        if (crc16(be, 4) != check) return false;
        value = data;
        return true;
    }

    // Gps::parse (gps.cpp:7-105)
    static bool position(const uint8_t* d, float& latOut, float& lonOut) {
        for (int i = 0; i < 6; i++) {
            if ((d[i] & 0x0F) > 9) return false;
        }
        float la = (d[0] & 0x0F) * 10 + (d[1] & 0x0F) + (float) (d[2] & 0x0F) / 6 + (float) (d[3] & 0x0F) / 60 +
                   (float) (d[4] & 0x0F) / 600 + (float) (d[5] & 0x0F) / 6000;
        uint8_t dir = d[3] & 0xF0;
        if (dir == 0x30) la *= -1;
        else if (dir != 0x50) return false;
        float lo = 0;   // uninitialised in the reference when neither branch below matches
        uint8_t b = d[4] & 0xF0;
        const uint8_t c = d[6];
        if (b == 0x50) {
            if (c >= 0x76 && c < 0x7f) lo = c - 0x76;
            else if (c >= 0x6c && c < 0x75) lo = 100 + (c - 0x6c);
            else if (c >= 0x26 && c < 0x6b) lo = 110 + (c - 0x26);
            else return false;
        } else if (b == 0x30) {
            if (c >= 0x26 && c < 0x7f) lo = 10 + (c - 0x26);
            else return false;
        }
        b = d[7];
        if (b > 0x58 && b <= 0x61) lo += (float) (b - 0x58) / 60;
        else if (b >= 0x26 && b <= 0x57) lo += (float) (10 + (b - 0x26)) / 60;
        else return false;
        b = d[8];
        if (b >= 0x1c && b < 0x7f) lo += (float) (b - 0x1c) / 6000;
        else return false;
        dir = d[5] & 0xF0;
        if (dir == 0x50) lo *= -1;
        else if (dir != 0x30) return false;
        if (la > 90 || la < -90 || lo > 180 || lo < -180) return false;
        latOut = la;
        lonOut = lo;
        return true;
    }

    // FramePhase::decodeV2DataChannel (ysf_phase.cpp:258-306)
    void dataChannelV2(const uint8_t* payload, unsigned frameNumber) {
        uint8_t dibits[100], packed[25], whitened[13], dch[13];
        for (int i = 0; i < 100; i++) dibits[i] = payload[(i % 5) * 72 + (i * 2) / 10];
        packDibits(dibits, 100, packed);
        viterbi(packed, 100, whitened);
        if (crc16(whitened, 10) != (uint16_t) (whitened[10] << 8 | whitened[11])) return;
        dewhiten(whitened, dch, 100);
        if (frameNumber < 6) {
            switch (frameNumber) {
                case 0: set(target, callsign(dch)); break;
                case 1: set(source, callsign(dch)); break;
                case 2: set(down, callsign(dch)); break;
                case 3: set(up, callsign(dch)); break;
            }
            dtNext = 0;
        }
        if (frameNumber >= 6 && frameNumber < 8) {
            const unsigned offset = frameNumber - 6;
            if (offset != dtNext) {
                dtNext = 0;
            } else {
                dtNext = offset + 1;
                std::memcpy(dt + offset * 10, dch, 10);
            }
        }
        if (dtNext >= 2) {
            if (dt[18] != 0x03) return;
            uint8_t sum = 0;
            for (int i = 0; i < 19; i++) sum = (uint8_t) (sum + dt[i]);
            if (sum != dt[19]) return;
            const uint32_t command = (uint32_t) dt[1] << 16 | (uint32_t) dt[2] << 8 | dt[3];
            float la = 0, lo = 0;
            const bool valid = command == 0x22625f && position(dt + 5, la, lo);
            setPosition(valid, la, lo);
        }
    }

    // FramePhase::decodeHeaderDataChannel (ysf_phase.cpp:317-349)
    bool dataChannelHeader(const uint8_t* in, uint8_t* dch) {
        uint8_t dibits[180], packed[45], whitened[23];
        for (int i = 0; i < 180; i++) {
            const int sp = (i % 9) * 20 + i / 9;
            dibits[i] = in[(sp / 36) * 72 + sp % 36];
        }
        packDibits(dibits, 180, packed);
        viterbi(packed, 180, whitened);
        if (crc16(whitened, 20) != (uint16_t) (whitened[20] << 8 | whitened[21])) return false;
        dewhiten(whitened, dch, 160);
        return true;
    }

    void emit(const uint8_t* p, size_t n) { out->bytes.insert(out->bytes.end(), p, p + n); }

    // FramePhase::process (ysf_phase.cpp:45-172); false = back to sync search, frame not consumed
    bool frame(const uint8_t* f) {
        if (hamming_distance(f, kSync, 20) <= 3) {
            if (++syncCount > 12) syncCount = 12;
        } else if (--syncCount < 0) {
            resetMeta();
            return false;
        }
        uint32_t value = 0;
        const bool fresh = parseFich(f + 20, value);
        if (fresh) {
            haveFich = true;
            fich = value;
        }
        const uint8_t* payload = f + 120;
        if (!haveFich) return true;
        const unsigned frameType = (fich >> 30) & 3, dataType = (fich >> 8) & 3;
        if (frameType == 1) {
            if (dataType == 0) {
                set(mode, "V1");
                for (int i = 0; i < 5; i++) {
                    uint8_t block[10] = {(uint8_t) dataType};
                    const uint8_t* in = payload + 36 + i * 72;
                    for (int k = 0; k < 36; k++) block[1 + k / 4] = (uint8_t) ((in[k] & 3) << (6 - 2 * (k % 4)));   // `=`, sic
                    emit(block, 10);
                }
            } else if (dataType == 2) {
                set(mode, "DN");
                for (int i = 0; i < 5; i++) {
                    const uint8_t* in = payload + 20 + i * 72;
                    uint8_t interleaved[13], plain[13] = {0}, clear[13];
                    packDibits(in, 52, interleaved);
                    for (int k = 0; k < 104; k++) {
                        const int o = (k * 4) % 104 + k * 4 / 104;
                        plain[k / 8] |= (uint8_t) (((interleaved[o / 8] >> (7 - o % 8)) & 1) << (7 - k % 8));
                    }
                    dewhiten(plain, clear, 104);
                    auto bitAt = [&](int i2) { return (clear[i2 / 8] >> (7 - i2 % 8)) & 1; };
                    uint8_t voice[49];
                    for (int t = 0; t < 27; t++) voice[t] = (uint8_t) (bitAt(3 * t) + bitAt(3 * t + 1) + bitAt(3 * t + 2) >= 2);
                    for (int k = 0; k < 22; k++) voice[27 + k] = (uint8_t) bitAt(81 + k);
                    // v2_voice_mapping (ysf_phase.hpp:45-51): three interleaved runs of 18, 18 and 13 positions
                    uint8_t block[8] = {(uint8_t) dataType};
                    for (int b = 0; b < 49; b++) {
                        int dst;
                        if (b < 18) dst = b < 14 ? 3 * b : 41 + 2 * (b - 14);
                        else if (b < 36) dst = b - 18 < 14 ? 3 * (b - 18) + 1 : 42 + 2 * (b - 32);
                        else dst = 3 * (b - 36) + 2;
                        block[1 + dst / 8] |= (uint8_t) (voice[b] << (7 - dst % 8));
                    }
                    emit(block, 8);
                }
                if (fresh) dataChannelV2(payload, (value >> 19) & 7);
            } else if (dataType == 3) {
                set(mode, "VW");
                const int first = expectSubFrame ? 3 : 0;
                expectSubFrame = false;
                for (int i = first; i < 5; i++) {
                    uint8_t block[19] = {(uint8_t) dataType};
                    packDibits(payload + i * 72, 72, block + 1);
                    emit(block, 19);
                }
            } else {
                set(mode, "FR data");
            }
        } else if (frameType == 0) {
            resetMeta();
            held++;
            uint8_t dch[20];
            if (dataChannelHeader(payload, dch)) {
                set(target, callsign(dch));
                set(source, callsign(dch + 10));
            }
            if (dataChannelHeader(payload + 36, dch)) {
                set(down, callsign(dch));
                set(up, callsign(dch + 10));
            }
            release();
            expectSubFrame = true;
        } else if (frameType == 2) {
            resetMeta();
        }
        return true;
    }
};

}  // namespace

void decode_ysf(const uint8_t* sym, size_t n, Decoded& out) {
    Ysf y;
    y.out = &out;
    size_t pos = 0;
    for (;;) {
        if (!y.framing) {
            if (n - pos <= 20) break;
            if (hamming_distance(sym + pos, kSync, 20) <= 3) {
                y.framing = true;
                y.syncCount = 0;
                y.haveFich = false;
                y.expectSubFrame = false;
                y.dtNext = 0;
            } else {
                pos++;
            }
        } else {
            if (n - pos <= 480) break;
            if (y.frame(sym + pos)) pos += 480;
            else y.framing = false;
        }
    }
}

}  // namespace port
