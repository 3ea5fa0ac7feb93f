// port_dmr.cpp — CPU restatement of the reference's DMR decoder incl. its metadata plane.
// TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
//
// Follows Digiham::Dmr::{SyncPhase,FramePhase} (reference src/dmr_decoder/dmr_phase.cpp:18-345), Cach/Tact
// (cach.cpp:11-31, tact.cpp:9-26), Emb (emb.cpp:9-24), EmbeddedCollector (embedded.cpp:20-94), SlotType
// (slottype.cpp:9-22), Lc (lc.cpp:8-43), MetaCollector/Slot (dmr_meta.cpp:7-179), TalkerAliasCollector
// (talkeralias.cpp:14-144) and Gps (gps.cpp:7-17), as one sequential walk over the symbol stream.  The decoder is
// driven like `while (canProcess()) process();` with the whole stream buffered, which gives the same result as any
// chunking (checked against the compiled reference in tests/test_oracle_cpu.py).
#include "port.hpp"

#include <cstring>

namespace port {

namespace {

const uint8_t kSyncBsData[24] = {3, 1, 3, 3, 3, 3, 1, 1, 1, 3, 3, 1, 1, 3, 1, 1, 3, 1, 3, 3, 1, 1, 3, 1};
const uint8_t kSyncBsVoice[24] = {1, 3, 1, 1, 1, 1, 3, 3, 3, 1, 1, 3, 3, 1, 3, 3, 1, 3, 1, 1, 3, 3, 1, 3};
const uint8_t kSyncMsData[24] = {3, 1, 1, 1, 3, 1, 1, 3, 3, 3, 1, 3, 1, 3, 3, 3, 3, 1, 1, 3, 1, 1, 1, 3};
const uint8_t kSyncMsVoice[24] = {1, 3, 3, 3, 1, 3, 3, 1, 1, 1, 3, 1, 3, 1, 1, 1, 1, 3, 3, 1, 3, 3, 3, 1};

enum { kData = 1, kVoice = 2 };

int syncType(const uint8_t* p) {
    if (hamming_distance(p, kSyncBsData, 24) <= 3) return kData;
    if (hamming_distance(p, kSyncBsVoice, 24) <= 3) return kVoice;
    if (hamming_distance(p, kSyncMsData, 24) <= 3) return kData;
    if (hamming_distance(p, kSyncMsVoice, 24) <= 3) return kVoice;
    return -1;
}

struct AliasCollector {
    uint8_t data[28] = {0};
    unsigned blocks = 0;

    unsigned bytes() const {
        int i = 0;
        for (; i < 4; i++) {
            const unsigned mask = (1u << (i + 1)) - 1;
            if ((blocks & mask) != mask) break;
        }
        return (unsigned) i * 7;
    }
    unsigned format() const { return data[0] >> 6; }
    unsigned length() const { return (data[0] & 0x3E) >> 1; }
    std::string text() const {
        if (!(blocks & 1)) return "";
        const unsigned n = bytes();
        std::string r;
        switch (format()) {
            case 0: {
                std::string all;
                for (unsigned i = 0; i < n; i += 7) {
                    const uint8_t* s = data + i;
                    const uint8_t ch[8] = {(uint8_t) (s[0] >> 1),
                                           (uint8_t) ((s[0] & 1) << 6 | s[1] >> 2),
                                           (uint8_t) ((s[1] & 3) << 5 | s[2] >> 3),
                                           (uint8_t) ((s[2] & 7) << 4 | s[3] >> 4),
                                           (uint8_t) ((s[3] & 15) << 3 | s[4] >> 5),
                                           (uint8_t) ((s[4] & 31) << 2 | s[5] >> 6),
                                           (uint8_t) ((s[5] & 63) << 1 | s[6] >> 7),
                                           (uint8_t) (s[6] & 127)};
                    all.append((const char*) ch, 8);
                }
                r = all.substr(1);
                break;
            }
            case 1: r = latin1_to_utf8(data + 1, n - 1); break;
            case 2: r.assign((const char*) data + 1, n - 1); break;
            case 3: {
                for (unsigned k = 0; k < (n - 1) / 2; k++) {
                    uint32_t u = (uint32_t) data[1 + 2 * k] << 8 | data[2 + 2 * k];
                    if (u >= 0xD800 && u < 0xDC00 && k + 1 < (n - 1) / 2) {
                        const uint32_t lo = (uint32_t) data[3 + 2 * k] << 8 | data[4 + 2 * k];
                        if (lo >= 0xDC00 && lo < 0xE000) {
                            u = 0x10000 + ((u - 0xD800) << 10) + (lo - 0xDC00);
                            k++;
                        }
                    }
                    if (u < 0x80) r += (char) u;
                    else if (u < 0x800) { r += (char) (0xC0 | u >> 6); r += (char) (0x80 | (u & 63)); }
                    else if (u < 0x10000) { r += (char) (0xE0 | u >> 12); r += (char) (0x80 | ((u >> 6) & 63)); r += (char) (0x80 | (u & 63)); }
                    else { r += (char) (0xF0 | u >> 18); r += (char) (0x80 | ((u >> 12) & 63)); r += (char) (0x80 | ((u >> 6) & 63)); r += (char) (0x80 | (u & 63)); }
                }
                break;
            }
        }
        if (r.size() > length()) r.resize(length());
        return r;
    }
    bool complete() const {
        if (!(blocks & 1)) return false;
        const int n = (int) bytes();
        switch (format()) {
            case 0: return (n * 7) / 8 - 1 >= (int) length();
            case 1: return n - 1 >= (int) length();
            case 2: return text().size() >= length();
            case 3: return (n - 1) / 2 >= (int) length();
        }
        return false;
    }
};

struct SlotMeta {
    bool dirty = false;
    int sync = -1, type = -1;
    uint32_t source = 0, target = 0;
    std::string alias;
    bool located = false;
    float lat = 0, lon = 0;

    template <typename T>
    void set(T& field, const T& v) {
        if (!(field == v)) {
            field = v;
            dirty = true;
        }
    }
    void softReset() {
        set(type, -1);
        set(source, 0u);
        set(target, 0u);
        set(alias, std::string());
        if (located) {
            located = false;
            dirty = true;
        }
    }
    void reset() {
        softReset();
        set(sync, -1);
    }
};

struct Dmr {
    // decoder-level
    bool framing = false;
    int slotFilter = 3;
    SlotMeta meta[2];
    Decoded* out = nullptr;
    // FramePhase members (dmr_phase.hpp:51-60)
    int syncCount = 0, slot = -1, stability = 0;
    int syncTypes[2] = {-1, -1};
    int slotSync[2] = {0, 0};
    int activeSlot = -1;
    unsigned superframe[2] = {0, 0};
    uint8_t embData[2][16];
    unsigned embFill[2] = {0, 0};
    AliasCollector alias[2];

    void startFraming() {
        framing = true;
        syncCount = 0;
        slot = -1;
        stability = 0;
        syncTypes[0] = syncTypes[1] = -1;
        slotSync[0] = slotSync[1] = 0;
        activeSlot = -1;
        superframe[0] = superframe[1] = 0;
        std::memset(embData, 0, sizeof(embData));
        embFill[0] = embFill[1] = 0;
        alias[0] = AliasCollector();
        alias[1] = AliasCollector();
    }

    // MetaCollector::sendMetaDataForSlot (dmr_meta.cpp:160-172)
    void publish(int s) {
        SlotMeta& m = meta[s];
        if (!m.dirty) return;
        std::map<std::string, std::string> kv;
        kv["protocol"] = "DMR";
        kv["slot"] = std::to_string(s);
        if (m.sync > 0) kv["sync"] = m.sync == kData ? "data" : m.sync == kVoice ? "voice" : "unknown";
        if (m.type > 0) kv["type"] = m.type == 1 ? "direct" : m.type == 2 ? "group" : "unknown";
        if (m.source > 0) kv["source"] = std::to_string(m.source);
        if (m.target > 0) kv["target"] = std::to_string(m.target);
        if (!m.alias.empty()) kv["talkeralias"] = m.alias;
        if (m.located) {
            kv["lat"] = std::to_string(m.lat);
            kv["lon"] = std::to_string(m.lon);
        }
        out->meta += serialize(kv);
        m.dirty = false;
    }
    void resetSlot(int s) {
        meta[s].reset();
        publish(s);
    }
    void resetAll() {
        meta[0].reset();
        meta[1].reset();
        publish(0);
        publish(1);
    }

    // FramePhase::handleLc (dmr_phase.cpp:304-339)
    void linkControl(int s, const uint8_t* lc) {
        const unsigned opcode = lc[0] & 0x3F;
        if (opcode == 0 || opcode == 3) {
            meta[s].set(meta[s].type, opcode == 0 ? 2 : 1);
            meta[s].set(meta[s].target, (uint32_t) lc[3] << 16 | (uint32_t) lc[4] << 8 | lc[5]);
            meta[s].set(meta[s].source, (uint32_t) lc[6] << 16 | (uint32_t) lc[7] << 8 | lc[8]);
            publish(s);
        } else if (opcode >= 4 && opcode <= 7) {
            std::memcpy(alias[s].data + (opcode - 4) * 7, lc + 2, 7);
            alias[s].blocks |= 1u << (opcode - 4);
            if (alias[s].complete()) {
                std::string a = alias[s].text();
                const size_t end = a.find_last_not_of('\0');
                a = end == std::string::npos ? "" : a.substr(0, end + 1);
                meta[s].set(meta[s].alias, a);
                publish(s);
            }
        } else if (opcode == 8) {
            const uint8_t* d = lc + 2;
            int32_t la = ((d[4] & 0x7F) << 16) | (d[5] << 8) | d[6];
            if (d[4] & 0x80) la *= -1;
            int32_t lo = (d[1] << 16) | (d[2] << 8) | d[3];
            if (d[0] & 1) lo *= -1;
            const float lat = 180.0f / (float) (1 << 24) * (float) la;
            const float lon = 360.0f / (float) (1 << 25) * (float) lo;
            SlotMeta& m = meta[s];
            if (!(m.located && m.lat == lat && m.lon == lon)) {
                m.located = true;
                m.lat = lat;
                m.lon = lon;
                m.dirty = true;
            }
            publish(s);
        }
    }

    // EmbeddedCollector::getLc (embedded.cpp:32-94)
    bool embeddedLc(int s, uint8_t* lc) {
        if (embFill[s] < 3) return false;
        uint32_t row[8] = {0};
        for (int i = 0; i < 16; i++) {
            for (int k = 0; k < 8; k++) row[k] = ((row[k] << 1) | ((embData[s][i] >> (7 - k)) & 1u)) & 0xFFFFu;
        }
        for (int k = 0; k < 7; k++) {
            if (!correct(H16_11, row[k])) return false;
        }
        uint32_t parity = 0;
        for (int k = 0; k < 8; k++) parity ^= row[k];
        if (parity) return false;
        // 77 payload bits: rows 0-1 carry 11, rows 2-6 carry 10 + one checksum bit
        std::memset(lc, 0, 9);
        int bit = 0;
        unsigned received = 0;
        for (int k = 0; k < 7; k++) {
            const int nbits = k < 2 ? 11 : 10;
            for (int b = 0; b < nbits; b++, bit++) lc[bit / 8] |= (uint8_t) (((row[k] >> (15 - b)) & 1u) << (7 - bit % 8));
            if (k >= 2) received |= ((row[k] >> 5) & 1u) << (6 - k);
        }
        unsigned sum = 0;
        for (int i = 0; i < 9; i++) sum += lc[i];
        return sum % 31 == received;
    }

    // counts a missing sync; true when the decoder has to fall back to sync search (dmr_phase.cpp:173-186,193-205)
    bool missedSync() {
        if (--slotSync[slot] < 0) {
            slotSync[slot] = 0;
            syncTypes[slot] = -1;
            resetSlot(slot);
            if (activeSlot == slot) activeSlot = -1;
        }
        if (--syncCount < 0) {
            resetAll();
            return true;
        }
        return false;
    }

    // FramePhase::process (dmr_phase.cpp:65-302); false = back to sync search, frame not consumed
    bool frame(const uint8_t* f) {
        // TACT: CACH bits 0,4,8,12,14,18,22 (cach.cpp:7,13-19), Hamming(7,4) always corrects
        static const int tactBits[7] = {0, 4, 8, 12, 14, 18, 22};
        uint32_t tact = 0;
        for (int b : tactBits) tact = (tact << 1) | ((f[b / 2] >> (1 - b % 2)) & 1u);
        correct(H7_4, tact);
        const int tc = (tact >> 5) & 1;
        const unsigned char next = (unsigned char) (slot ^ 1);
        if (tc != next) {
            if (stability < 5) {
                stability = 0;
                slot = tc;
                const int other = slot ^ 1;
                syncTypes[other] = -1;
                resetSlot(other);
                if (activeSlot == other) activeSlot = -1;
            } else {
                stability--;
                if (slot != -1) slot = next;
            }
        } else {
            if (++stability > 100) stability = 100;
            slot = next;
        }

        const int st = syncType(f + 66);
        if (st > 0) {
            if (++syncCount > 5) syncCount = 5;
            if (++slotSync[slot] > 5) slotSync[slot] = 5;
            const bool soft = syncTypes[slot] == kVoice && st != syncTypes[slot];
            syncTypes[slot] = st;
            meta[slot].set(meta[slot].sync, st);
            if (soft) meta[slot].softReset();
            publish(slot);
            superframe[slot] = 0;
            embFill[slot] = 0;
        } else if (syncTypes[slot] == kVoice && superframe[slot] < 5) {
            superframe[slot]++;
            uint32_t emb = 0;
            for (int i = 0; i < 4; i++) emb = (emb << 2) | f[66 + i];
            for (int i = 0; i < 4; i++) emb = (emb << 2) | f[86 + i];
            emb &= 0xFFFF;
            if (correct(QR16_7, emb)) {
                if (++syncCount > 5) syncCount = 5;
                if (++slotSync[slot] > 5) slotSync[slot] = 5;
                uint8_t fragment[4] = {0, 0, 0, 0};
                for (int i = 0; i < 16; i++) fragment[i / 4] |= (uint8_t) (f[70 + i] << (6 - (i % 4) * 2));
                const unsigned lcss = (emb >> 9) & 3;
                auto collect = [&]() {
                    if (embFill[slot] > 3) return;
                    std::memcpy(embData[slot] + embFill[slot] * 4, fragment, 4);
                    embFill[slot]++;
                };
                if (lcss == 1) {            // first fragment
                    embFill[slot] = 0;
                    collect();
                } else if (lcss == 3) {     // continuation
                    collect();
                } else if (lcss == 2) {     // last fragment
                    collect();
                    uint8_t lc[9];
                    if (embeddedLc(slot, lc)) linkControl(slot, lc);
                    embFill[slot] = 0;
                }
            } else if (missedSync()) {
                return false;
            }
        } else {
            superframe[slot] = 0;
            embFill[slot] = 0;
            if (missedSync()) return false;
        }

        if (syncTypes[slot] == kVoice) {
            if (((slot + 1) & slotFilter) && (activeSlot == -1 || activeSlot == slot)) {
                activeSlot = slot;
                uint8_t payload[27] = {0};
                for (int i = 0; i < 108; i++) {
                    const uint8_t d = f[i < 54 ? 12 + i : 90 + (i - 54)] & 3;
                    payload[i / 4] |= (uint8_t) (d << (6 - 2 * (i % 4)));
                }
                out->bytes.insert(out->bytes.end(), payload, payload + 27);
            }
        } else {
            if (activeSlot == slot) activeSlot = -1;
            alias[slot].blocks = 0;
            if (syncTypes[slot] == kData) {
                uint32_t slotType = 0;
                for (int i = 0; i < 5; i++) slotType = (slotType << 2) | (f[61 + i] & 3u);
                for (int i = 0; i < 5; i++) slotType = (slotType << 2) | (f[90 + i] & 3u);
                if (correct(GOLAY20_8, slotType)) {
                    const unsigned dataType = (slotType >> 12) & 15;
                    if (dataType != 8) {
                        uint8_t payload[25] = {0};
                        for (int k = 0; k < 98; k++) {
                            const uint8_t d = f[k < 49 ? 12 + k : 95 + (k - 49)] & 3;
                            payload[k / 4] |= (uint8_t) (d << (6 - 2 * (k % 4)));
                        }
                        uint8_t lc[12];
                        if (bptc_196_96(payload, lc)) {
                            if (dataType == 1) {
                                linkControl(slot, lc);
                            } else if (dataType == 2 || dataType == 9) {
                                meta[slot].softReset();
                                publish(slot);
                            }
                        }
                    }
                }
            } else {
                resetSlot(slot);
            }
        }
        return true;
    }
};

}  // namespace

void decode_dmr(const uint8_t* sym, size_t n, int slotFilter, Decoded& out) {
    Dmr d;
    d.out = &out;
    d.slotFilter = slotFilter;
    size_t pos = 0;
    for (;;) {
        if (!d.framing) {
            if (n - pos <= 90) break;                 // SyncPhase needs SYNC_SIZE + syncOffset buffered
            if (syncType(sym + pos + 66) > 0) d.startFraming();
            else pos++;
        } else {
            if (n - pos <= 144) break;
            if (d.frame(sym + pos)) pos += 144;
            else d.framing = false;
        }
    }
}

}  // namespace port
