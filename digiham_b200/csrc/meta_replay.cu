// meta_replay.cu — host side of the metadata plane: turns the GPU's event records back into the reference's
// `key:value;key:value\n` lines (StringSerializer, reference src/lib/meta.cpp:8-17; keys come out sorted because
// the reference collects them in a std::map).
//
// DMR: restates the observable behaviour of Dmr::Slot / Dmr::MetaCollector (src/dmr_decoder/dmr_meta.cpp:7-179),
// FramePhase::handleLc (dmr_phase.cpp:304-339), TalkerAliasCollector (talkeralias.cpp:14-144) and Dmr::Gps
// (gps.cpp:7-17).  This is string formatting driven by rare events and deliberately stays on the host.
#include "meta_replay.hpp"

#include <cstdio>
#include <cstring>
#include <map>

namespace dh {

namespace {

std::string serialize(const std::map<std::string, std::string>& kv) {
    std::string out;
    bool first = true;
    for (const auto& it : kv) {
        if (!first) out += ';';
        first = false;
        out += it.first;
        out += ':';
        out += it.second;
    }
    out += '\n';
    return out;
}

// ISO-8859-1 -> UTF-8; like the reference's ICU call (src/lib/charset.cpp:16-24) the result ends at the first NUL
std::string latin1_to_utf8(const unsigned char* p, size_t n) {
    std::string r;
    for (size_t i = 0; i < n; i++) {
        const unsigned char ch = p[i];
        if (ch == 0) break;
        if (ch < 0x80) {
            r.push_back((char) ch);
        } else {
            r.push_back((char) (0xC0 | (ch >> 6)));
            r.push_back((char) (0x80 | (ch & 0x3F)));
        }
    }
    return r;
}

void append_utf8(std::string& r, uint32_t cp) {
    if (cp < 0x80) {
        r.push_back((char) cp);
    } else if (cp < 0x800) {
        r.push_back((char) (0xC0 | (cp >> 6)));
        r.push_back((char) (0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
        r.push_back((char) (0xE0 | (cp >> 12)));
        r.push_back((char) (0x80 | ((cp >> 6) & 0x3F)));
        r.push_back((char) (0x80 | (cp & 0x3F)));
    } else {
        r.push_back((char) (0xF0 | (cp >> 18)));
        r.push_back((char) (0x80 | ((cp >> 12) & 0x3F)));
        r.push_back((char) (0x80 | ((cp >> 6) & 0x3F)));
        r.push_back((char) (0x80 | (cp & 0x3F)));
    }
}

// ---- DMR ---------------------------------------------------------------------------------------------------------

struct TalkerAlias {
    unsigned char data[28] = {0};
    unsigned blocks = 0;

    void reset() { blocks = 0; }
    void setBlock(int block, const unsigned char* src) {
        std::memcpy(data + block * 7, src, 7);
        blocks |= 1u << block;
    }
    bool hasHeader() const { return blocks & 1u; }
    unsigned format() const { return data[0] >> 6; }
    unsigned length() const { return (data[0] & 0x3E) >> 1; }
    unsigned collectedBytes() const {
        int i;
        for (i = 0; i < 4; i++) {
            const unsigned mask = (1u << (i + 1)) - 1;
            if ((blocks & mask) != mask) break;
        }
        return (unsigned) i * 7;
    }
    std::string contents() const {
        if (!hasHeader()) return "";
        const unsigned bytes = collectedBytes();
        std::string result;
        switch (format()) {
            case 0: {   // 7 bit: eight characters per seven bytes; the first one is made of header bits
                std::string all;
                for (unsigned i = 0; i < bytes; i += 7) {
                    const unsigned char* s = data + i;
                    unsigned long long v = 0;
                    for (int k = 0; k < 7; k++) v = (v << 8) | s[k];
                    for (int k = 0; k < 8; k++) all.push_back((char) ((v >> (49 - 7 * k)) & 0x7F));
                }
                result = all.substr(1);
                break;
            }
            case 1:
                result = latin1_to_utf8(data + 1, bytes - 1);
                break;
            case 2:
                result = std::string((const char*) data + 1, bytes - 1);
                break;
            case 3: {
                const unsigned chars = (bytes - 1) / 2;
                const unsigned char* src = data + 1;
                for (unsigned k = 0; k < chars; k++) {
                    uint32_t u = (uint32_t) (src[k * 2] << 8) | src[k * 2 + 1];
                    if (u >= 0xD800 && u < 0xDC00 && k + 1 < chars) {
                        const uint32_t lo = (uint32_t) (src[k * 2 + 2] << 8) | src[k * 2 + 3];
                        if (lo >= 0xDC00 && lo < 0xE000) {
                            u = 0x10000 + ((u - 0xD800) << 10) + (lo - 0xDC00);
                            k++;
                        }
                    }
                    append_utf8(result, u);
                }
                break;
            }
        }
        if (result.length() > length()) result = result.substr(0, length());
        return result;
    }
    bool complete() const {
        if (!hasHeader()) return false;
        const int bytes = (int) collectedBytes();
        switch (format()) {
            case 0: return ((bytes * 7) / 8) - 1 >= (int) length();
            case 1: return bytes - 1 >= (int) length();
            case 2: return contents().length() >= length();
            case 3: return (bytes - 1) / 2 >= (int) length();
        }
        return false;
    }
};

struct DmrSlot {
    bool dirty = false;
    int sync = -1;
    int type = -1;
    uint32_t source = 0;
    uint32_t target = 0;
    std::string alias;
    bool hasCoord = false;
    float lat = 0, lon = 0;

    void setSync(int v) { if (sync != v) { sync = v; dirty = true; } }
    void setType(int v) { if (type != v) { type = v; dirty = true; } }
    void setSource(uint32_t v) { if (source != v) { source = v; dirty = true; } }
    void setTarget(uint32_t v) { if (target != v) { target = v; dirty = true; } }
    void setAlias(const std::string& v) { if (alias != v) { alias = v; dirty = true; } }
    void clearCoord() { if (hasCoord) { hasCoord = false; dirty = true; } }
    void setCoord(float la, float lo) {
        if (hasCoord && lat == la && lon == lo) return;
        hasCoord = true;
        lat = la;
        lon = lo;
        dirty = true;
    }
    void softReset() {
        setType(-1);
        setSource(0);
        setTarget(0);
        setAlias("");
        clearCoord();
    }
    void reset() {
        softReset();
        setSync(-1);
    }
};

class DmrReplay: public MetaReplay {
    public:
        void apply(const DecEvent* ev, uint32_t n, std::string& out) override {
            for (uint32_t i = 0; i < n; i++) {
                const DecEvent& e = ev[i];
                const int s = e.slot & 1;
                switch (e.kind) {
                    case 1: slots[s].reset(); break;
                    case 2:
                        slots[s].setSync(e.a);
                        if (e.b) slots[s].softReset();
                        break;
                    case 3: slots[s].softReset(); break;
                    case 4: handleLc(s, e.data); break;
                    case 5: ta[s].reset(); continue;   // no withSlot involved
                    default: continue;
                }
                flush(s, out);
            }
        }
    private:
        DmrSlot slots[2];
        TalkerAlias ta[2];

        void handleLc(int s, const uint8_t* lc) {
            const unsigned opcode = lc[0] & 0x3F;
            switch (opcode) {
                case 0:
                case 3:
                    slots[s].setType(opcode == 0 ? 2 : 1);
                    slots[s].setTarget((uint32_t) lc[3] << 16 | (uint32_t) lc[4] << 8 | lc[5]);
                    slots[s].setSource((uint32_t) lc[6] << 16 | (uint32_t) lc[7] << 8 | lc[8]);
                    break;
                case 4: case 5: case 6: case 7:
                    ta[s].setBlock((int) opcode - 4, lc + 2);
                    if (ta[s].complete()) {
                        std::string alias = ta[s].contents();
                        const size_t end = alias.find_last_not_of('\0');
                        alias = end == std::string::npos ? "" : alias.substr(0, end + 1);
                        slots[s].setAlias(alias);
                    }
                    break;
                case 8: {
                    const uint8_t* d = lc + 2;
                    int32_t latBits = ((d[4] & 0x7F) << 16) | (d[5] << 8) | d[6];
                    if (d[4] & 0x80) latBits *= -1;
                    int32_t lonBits = (d[1] << 16) | (d[2] << 8) | d[3];
                    if (d[0] & 0x01) lonBits *= -1;
                    const float la = 180.0f / (float) (1 << 24) * (float) latBits;
                    const float lo = 360.0f / (float) (1 << 25) * (float) lonBits;
                    slots[s].setCoord(la, lo);
                    break;
                }
            }
        }

        // MetaCollector::sendMetaDataForSlot (dmr_meta.cpp:160-172)
        void flush(int s, std::string& out) {
            DmrSlot& sl = slots[s];
            if (!sl.dirty) return;
            std::map<std::string, std::string> kv;
            kv["protocol"] = "DMR";
            kv["slot"] = std::to_string(s);
            if (sl.sync > 0) kv["sync"] = sl.sync == 1 ? "data" : (sl.sync == 2 ? "voice" : "unknown");
            if (sl.type > 0) kv["type"] = sl.type == 1 ? "direct" : (sl.type == 2 ? "group" : "unknown");
            if (sl.source > 0) kv["source"] = std::to_string(sl.source);
            if (sl.target > 0) kv["target"] = std::to_string(sl.target);
            if (!sl.alias.empty()) kv["talkeralias"] = sl.alias;
            if (sl.hasCoord) {
                kv["lat"] = std::to_string(sl.lat);
                kv["lon"] = std::to_string(sl.lon);
            }
            out += serialize(kv);
            sl.dirty = false;
        }
};

}  // namespace

MetaReplay* make_dmr_replay() { return new DmrReplay(); }

}  // namespace dh
