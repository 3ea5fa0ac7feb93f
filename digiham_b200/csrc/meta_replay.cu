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
#include <vector>
#include <cstdlib>

namespace dh {

namespace {
std::string serialize(const std::map<std::string, std::string>& kv);
}

void MetaReplay::emit(const std::map<std::string, std::string>& kv, std::string& out) {
    out += serialize(kv);
    emit_kv_only(kv);
}

void MetaReplay::emit_kv_only(const std::map<std::string, std::string>& kv) {
    if (kv_sink) {
        auto put16 = [&](size_t v) {
            kv_sink->push_back((char) (v & 0xFF));
            kv_sink->push_back((char) ((v >> 8) & 0xFF));
        };
        put16(kv.size());
        for (const auto& it : kv) {
            put16(it.first.size());
            kv_sink->append(it.first);
            put16(it.second.size());
            kv_sink->append(it.second);
        }
    }
}

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
    // bytes available without a gap: 7 per block, counting the blocks received contiguously from block 0
    unsigned collectedBytes() const {
        const unsigned leading = (unsigned) __builtin_ctz(~blocks & 0x1Fu);   // trailing one-bits of the block mask
        return (leading < 4 ? leading : 4) * 7;
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
        void save(std::string& blob) const override {
            StateWriter w{blob};
            const_cast<DmrReplay*>(this)->fields(w);
        }
        bool load(const uint8_t* data, size_t len) override {
            StateReader r{data, data + len};
            fields(r);
            return r.ok && r.p == r.end;
        }
    private:
        DmrSlot slots[2];
        TalkerAlias ta[2];

        template <class A> void fields(A& a) {
            for (int s = 0; s < 2; s++) {
                a.val(slots[s].dirty);
                a.val(slots[s].sync);
                a.val(slots[s].type);
                a.val(slots[s].source);
                a.val(slots[s].target);
                a.str(slots[s].alias);
                a.val(slots[s].hasCoord);
                a.val(slots[s].lat);
                a.val(slots[s].lon);
                a.pod(ta[s].data, sizeof(ta[s].data));
                a.val(ta[s].blocks);
            }
        }

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
            // keys in std::map (sorted) order: lat, lon, protocol, slot, source, sync, talkeralias, target, type
            MetaLine l(out, kv_sink);
            if (sl.hasCoord) {
                l.add("lat", std::to_string(sl.lat));
                l.add("lon", std::to_string(sl.lon));
            }
            l.add("protocol", "DMR", 3);
            l.add("slot", s ? "1" : "0", 1);
            if (sl.source > 0) l.add_uint("source", sl.source);
            if (sl.sync > 0) l.add("sync", sl.sync == 1 ? "data" : (sl.sync == 2 ? "voice" : "unknown"));
            if (!sl.alias.empty()) l.add("talkeralias", sl.alias);
            if (sl.target > 0) l.add_uint("target", sl.target);
            if (sl.type > 0) l.add("type", sl.type == 1 ? "direct" : (sl.type == 2 ? "group" : "unknown"));
            l.finish();
            sl.dirty = false;
        }
};

// ---- YSF ---------------------------------------------------------------------------------------------------------
// Restates Ysf::MetaCollector (src/ysf_decoder/ysf_meta.cpp:7-105) incl. the hold/release batching of the base
// class (src/lib/meta.cpp:71-100), FramePhase::treatYsfString (ysf_phase.cpp:351-361), DataCollector / DataFrame
// (data.cpp:15-88) and Ysf::Gps::parse (gps.cpp:7-105).

class YsfReplay: public MetaReplay {
    public:
        void apply(const DecEvent* ev, uint32_t n, std::string& out) override {
            for (uint32_t i = 0; i < n; i++) {
                const DecEvent& e = ev[i];
                switch (e.kind) {
                    case 1: {
                        static const char* const names[] = {"", "V1", "DN", "VW", "FR data"};
                        set(mode, names[e.a <= 4 ? e.a : 0], out);
                        break;
                    }
                    case 2: reset(out); break;
                    case 3: held++; break;
                    case 4: {
                        const std::string v = treat(e.data);
                        switch (e.a) {
                            case 0: set(destination, v, out); break;
                            case 1: set(source, v, out); break;
                            case 2: set(down, v, out); break;
                            case 3: set(up, v, out); break;
                        }
                        break;
                    }
                    case 5: release(out); break;
                    case 6: dcNext = 0; break;
                    case 7:
                        if (e.a != dcNext) {
                            dcNext = 0;
                        } else {
                            dcNext = e.a + 1;
                            std::memcpy(dcData + (e.a & 1) * 10, e.data, 10);
                        }
                        break;
                    case 8:
                        if (dcNext >= 2) checkDataFrame(out);
                        break;
                }
            }
        }
        void save(std::string& blob) const override {
            StateWriter w{blob};
            const_cast<YsfReplay*>(this)->fields(w);
        }
        bool load(const uint8_t* data, size_t len) override {
            StateReader r{data, data + len};
            fields(r);
            return r.ok && r.p == r.end;
        }
    private:
        std::string mode, destination, source, up, down;
        bool hasCoord = false;
        float lat = 0, lon = 0;
        int held = 0;
        bool dirty = false;
        unsigned dcNext = 0;
        unsigned char dcData[20] = {0};

        template <class A> void fields(A& a) {
            a.str(mode);
            a.str(destination);
            a.str(source);
            a.str(up);
            a.str(down);
            a.val(hasCoord);
            a.val(lat);
            a.val(lon);
            a.val(held);
            a.val(dirty);
            a.val(dcNext);
            a.pod(dcData, sizeof(dcData));
        }

        void send(std::string& out) {
            if (held) {
                dirty = true;
                return;
            }
            // keys in sorted order: down, lat, lon, mode, protocol, source, target, up
            MetaLine l(out, kv_sink);
            if (!down.empty()) l.add("down", down);
            if (hasCoord) {
                l.add("lat", std::to_string(lat));
                l.add("lon", std::to_string(lon));
            }
            if (!mode.empty()) l.add("mode", mode);
            l.add("protocol", "YSF", 3);
            if (!source.empty()) l.add("source", source);
            if (!destination.empty()) l.add("target", destination);
            if (!up.empty()) l.add("up", up);
            l.finish();
        }
        void set(std::string& field, const std::string& v, std::string& out) {
            if (field == v) return;
            field = v;
            send(out);
        }
        void setGps(bool valid, float la, float lo, std::string& out) {
            if (!valid && !hasCoord) return;
            if (valid && hasCoord && lat == la && lon == lo) return;
            hasCoord = valid;
            lat = la;
            lon = lo;
            send(out);
        }
        void release(std::string& out) {
            held--;
            if (held == 0) {
                if (dirty) send(out);
                dirty = false;
            }
        }
        void reset(std::string& out) {
            held++;
            set(mode, "", out);
            set(destination, "", out);
            set(source, "", out);
            set(up, "", out);
            set(down, "", out);
            setGps(false, 0, 0, out);
            release(out);
        }
        static std::string treat(const uint8_t* input) {
            size_t length = 10;
            for (char ch : {'\n', ' '}) {
                const void* end = std::memchr(input, ch, length);
                if (end != nullptr) length = (size_t) ((const uint8_t*) end - input);
            }
            return latin1_to_utf8(input, length);
        }
        void checkDataFrame(std::string& out) {
            if (dcData[18] != 0x03) return;
            uint8_t checksum = 0;
            for (int i = 0; i < 19; i++) checksum = (uint8_t) (checksum + dcData[i]);
            if (checksum != dcData[19]) return;
            const uint32_t command = (uint32_t) dcData[1] << 16 | (uint32_t) dcData[2] << 8 | dcData[3];
            float la = 0, lo = 0;
            bool valid = command == 0x22625f && parseGps(dcData + 5, la, lo);   // COMMAND_SHORT_GPS
            setGps(valid, la, lo, out);
        }
        static bool parseGps(const uint8_t* d, float& latOut, float& lonOut) {
            for (int i = 0; i < 6; i++) {
                if ((d[i] & 0x0F) > 9) return false;
            }
            float la = (float) ((d[0] & 0x0F) * 10 + (d[1] & 0x0F));
            la = la + (float) (d[2] & 0x0F) / 6;
            la = la + (float) (d[3] & 0x0F) / 60;
            la = la + (float) (d[4] & 0x0F) / 600;
            la = la + (float) (d[5] & 0x0F) / 6000;
            uint8_t direction = d[3] & 0xF0;
            if (direction == 0x50) {
            } else if (direction == 0x30) {
                la *= -1;
            } else {
                return false;
            }
            float lo = 0;   // the reference leaves lon uninitialised when neither branch matches (gps.cpp:37-55)
            uint8_t b = d[4] & 0xF0;
            const uint8_t c = d[6];
            if (b == 0x50) {
                if (c >= 0x76 && c < 0x7f) lo = (float) (c - 0x76);
                else if (c >= 0x6c && c < 0x75) lo = (float) (100 + (c - 0x6c));
                else if (c >= 0x26 && c < 0x6b) lo = (float) (110 + (c - 0x26));
                else return false;
            } else if (b == 0x30) {
                if (c >= 0x26 && c < 0x7f) lo = (float) (10 + (c - 0x26));
                else return false;
            }
            b = d[7];
            if (b > 0x58 && b <= 0x61) lo += (float) (b - 0x58) / 60;
            else if (b >= 0x26 && b <= 0x57) lo += (float) (10 + (b - 0x26)) / 60;
            else return false;
            b = d[8];
            if (b >= 0x1c && b < 0x7f) lo += (float) (b - 0x1c) / 6000;
            else return false;
            direction = d[5] & 0xF0;
            if (direction == 0x50) lo *= -1;
            else if (direction != 0x30) return false;
            if (la > 90 || la < -90) return false;
            if (lo > 180 || lo < -180) return false;
            latOut = la;
            lonOut = lo;
            return true;
        }
};

// ---- NXDN --------------------------------------------------------------------------------------------------------
// Nxdn::MetaCollector (reference src/nxdn_decoder/nxdn_meta.cpp:6-76): every setter sends on its own unless held.
class NxdnReplay: public MetaReplay {
    public:
        void apply(const DecEvent* ev, uint32_t n, std::string& out) override {
            for (uint32_t i = 0; i < n; i++) {
                const DecEvent& e = ev[i];
                switch (e.kind) {
                    case 1: setStr(sync, "voice", out); break;
                    case 2:   // setFromSacch (nxdn_meta.cpp:54-66)
                        setStr(type, e.a == 1 ? "conference" : (e.a == 2 ? "individual" : ""), out);
                        setNum(source, (unsigned) e.data[0] << 8 | e.data[1], out);
                        setNum(destination, (unsigned) e.data[2] << 8 | e.data[3], out);
                        break;
                    case 3:   // reset (nxdn_meta.cpp:68-75)
                        held++;
                        setStr(sync, "", out);
                        setStr(type, "", out);
                        setNum(source, 0, out);
                        setNum(destination, 0, out);
                        if (--held == 0) {
                            if (dirty) send(out);
                            dirty = false;
                        }
                        break;
                }
            }
        }
        void save(std::string& blob) const override {
            StateWriter w{blob};
            const_cast<NxdnReplay*>(this)->fields(w);
        }
        bool load(const uint8_t* data, size_t len) override {
            StateReader r{data, data + len};
            fields(r);
            return r.ok && r.p == r.end;
        }
    private:
        std::string sync, type;
        unsigned source = 0, destination = 0;
        int held = 0;
        bool dirty = false;

        template <class A> void fields(A& a) {
            a.str(sync);
            a.str(type);
            a.val(source);
            a.val(destination);
            a.val(held);
            a.val(dirty);
        }

        void send(std::string& out) {
            if (held) {
                dirty = true;
                return;
            }
            // keys in sorted order: destination, protocol, source, sync, type
            MetaLine l(out, kv_sink);
            if (destination != 0) l.add_uint("destination", destination);
            l.add("protocol", "NXDN", 4);
            if (source != 0) l.add_uint("source", source);
            if (!sync.empty()) l.add("sync", sync);
            if (!type.empty()) l.add("type", type);
            l.finish();
        }
        void setStr(std::string& field, const std::string& v, std::string& out) {
            if (field == v) return;
            field = v;
            send(out);
        }
        void setNum(unsigned& field, unsigned v, std::string& out) {
            if (field == v) return;
            field = v;
            send(out);
        }
};

// ---- D-Star ------------------------------------------------------------------------------------------------------
// Host half of DStar::VoicePhase (reference src/dstar_decoder/dstar_phase.cpp:151-278: slow-data reassembly, DPRS
// and NMEA sentences), Header's callsign getters (header.cpp:150-178) and DStar::MetaCollector
// (dstar_meta.cpp:5-130).  Where the reference would throw or index out of range (std::stof on a non-numeric NMEA
// field, fewer than six GGA fields) the sentence is skipped.
class DstarReplay: public MetaReplay {
    public:
        void apply(const DecEvent* ev, uint32_t n, std::string& out) override {
            for (uint32_t i = 0; i < n; i++) {
                const DecEvent& e = ev[i];
                switch (e.kind) {
                    case 1:
                        if (e.a < 4) std::memcpy(radioHeader + 12 * e.a, e.data, e.a < 3 ? 12 : 5);
                        if (e.a == 3) setFromHeader(radioHeader, out);
                        break;
                    case 2:
                        resetFrames();
                        simpleData.clear();
                        break;
                    case 3: collect(e.data); break;
                    case 4:
                        if (e.a) set(sync, "voice", out);
                        parseFrameData(out);
                        resetFrames();
                        break;
                    case 5: reset(out); break;
                }
            }
        }
        void save(std::string& blob) const override {
            StateWriter w{blob};
            const_cast<DstarReplay*>(this)->fields(w);
        }
        bool load(const uint8_t* data, size_t len) override {
            StateReader r{data, data + len};
            fields(r);
            return r.ok && r.p == r.end;
        }
    private:
        std::string sync, message, departure, destination, ourCall, yourCall, dprs;
        bool located = false;
        float lat = 0, lon = 0;
        int held = 0;
        bool dirty = false;
        unsigned char radioHeader[41] = {0};
        unsigned char msg[20] = {0};
        unsigned msgBlocks = 0;
        unsigned char hdr[41] = {0};
        unsigned hdrCount = 0;
        std::string simpleData;

        template <class A> void fields(A& a) {
            a.str(sync);
            a.str(message);
            a.str(departure);
            a.str(destination);
            a.str(ourCall);
            a.str(yourCall);
            a.str(dprs);
            a.val(located);
            a.val(lat);
            a.val(lon);
            a.val(held);
            a.val(dirty);
            a.pod(radioHeader, sizeof(radioHeader));
            a.pod(msg, sizeof(msg));
            a.val(msgBlocks);
            a.pod(hdr, sizeof(hdr));
            a.val(hdrCount);
            a.str(simpleData);
        }

        void send(std::string& out) {
            if (held) {
                dirty = true;
                return;
            }
            // keys in sorted order: departure, destination, dprs, lat, lon, message, ourcall, protocol, sync, yourcall
            MetaLine l(out, kv_sink);
            if (!departure.empty()) l.add("departure", departure);
            if (!destination.empty()) l.add("destination", destination);
            if (!dprs.empty()) l.add("dprs", dprs);
            if (located) {
                l.add("lat", std::to_string(lat));
                l.add("lon", std::to_string(lon));
            }
            if (!message.empty()) l.add("message", message);
            if (!ourCall.empty()) l.add("ourcall", ourCall);
            l.add("protocol", "DSTAR", 5);
            if (!sync.empty()) l.add("sync", sync);
            if (!yourCall.empty()) l.add("yourcall", yourCall);
            l.finish();
        }
        void set(std::string& field, const std::string& v, std::string& out) {
            if (field == v) return;
            field = v;
            send(out);
        }
        void setGps(bool valid, float la, float lo, std::string& out) {
            if (!valid && !located) return;
            if (valid && located && lat == la && lon == lo) return;
            located = valid;
            lat = la;
            lon = lo;
            send(out);
        }
        void release(std::string& out) {
            if (--held == 0) {
                if (dirty) send(out);
                dirty = false;
            }
        }
        void reset(std::string& out) {
            held++;
            set(sync, "", out);
            set(message, "", out);
            set(departure, "", out);
            set(destination, "", out);
            set(ourCall, "", out);
            set(yourCall, "", out);
            set(dprs, "", out);
            setGps(false, 0, 0, out);
            release(out);
        }
        static std::string field(const unsigned char* p, size_t n) {
            std::string s = latin1_to_utf8(p, n);
            s.erase(s.find_last_not_of(' ') + 1);
            return s;
        }
        void setFromHeader(const unsigned char* h, std::string& out) {
            held++;
            set(sync, (h[0] >> 7) & 1 ? "data" : "voice", out);
            set(departure, field(h + 11, 8), out);
            set(destination, field(h + 3, 8), out);
            std::string own = field(h + 27, 8);
            const std::string suffix = field(h + 35, 4);
            if (!suffix.empty()) own += "/" + suffix;
            set(ourCall, own, out);
            set(yourCall, field(h + 19, 8), out);
            release(out);
        }
        void resetFrames() {
            std::memset(msg, 0, sizeof(msg));
            msgBlocks = 0;
            std::memset(hdr, 0, sizeof(hdr));
            hdrCount = 0;
        }
        void collect(const unsigned char* block) {
            const unsigned n = block[0] & 0x0F;
            switch (block[0] >> 4) {
                case 4:
                    if (n > 3) break;
                    std::memcpy(msg + n * 5, block + 1, 5);
                    msgBlocks |= 1u << n;
                    break;
                case 5:
                    if (n > 5 || hdrCount + n > 41) break;
                    std::memcpy(hdr + hdrCount, block + 1, n);
                    hdrCount += n;
                    break;
                case 3:
                    if (n > 5) break;
                    simpleData.append(reinterpret_cast<const char*>(block + 1), n);
                    break;
                default: break;
            }
        }
        static unsigned crcOf(const unsigned char* data, size_t len) {
            unsigned c = 0xFFFF;
            for (size_t k = 0; k < len; k++) {
                for (int i = 0; i < 8; i++) {
                    c ^= (data[k] >> i) & 1u;
                    c = (c & 1u) ? ((c >> 1) ^ 0x8408u) : (c >> 1);
                }
            }
            return c ^ 0xFFFFu;
        }
        void parseNmea(const std::string& input, std::string& out) {
            const size_t star = input.find_last_of('*');
            if (star == std::string::npos || star + 2 > input.length()) return;
            const std::string body = input.substr(1, star - 1);
            if (body.length() < 2) return;
            unsigned checksum = 0;
            for (char ch : body) checksum ^= (unsigned char) ch;
            if (checksum != (unsigned) std::strtoul(input.substr(star + 1, 2).c_str(), nullptr, 16)) return;
            std::vector<std::string> fields;
            size_t from = 0;
            while (from <= body.length()) {
                const size_t comma = body.find(',', from);
                if (comma == std::string::npos) {
                    if (from < body.length()) fields.push_back(body.substr(from));
                    break;
                }
                fields.push_back(body.substr(from, comma - from));
                from = comma + 1;
            }
            if (body.substr(2, 3) != "GGA" || fields.size() < 6) return;
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
            setGps(true, la, lo, out);
        }
        void parseFrameData(std::string& out) {
            if (msgBlocks == 0x0F) set(message, latin1_to_utf8(msg, 20), out);
            if (hdrCount == 41 && crcOf(hdr, 39) == ((unsigned) hdr[39] | (unsigned) hdr[40] << 8)) setFromHeader(hdr, out);
            size_t pos;
            while ((pos = simpleData.find('\r')) != std::string::npos) {
                const std::string s = simpleData.substr(0, pos + 1);
                if (s.length() >= 10 && s.compare(0, 5, "$$CRC") == 0 && s[9] == ',') {
                    const unsigned check = (unsigned) std::strtoul(s.substr(5, 4).c_str(), nullptr, 16);
                    if (crcOf(reinterpret_cast<const unsigned char*>(s.data()) + 10, s.length() - 10) == (check & 0xFFFFu))
                        set(dprs, s.substr(10, s.length() - 11), out);
                } else if (s.length() > 5 && s[0] == '$') {
                    parseNmea(s, out);
                }
                simpleData = simpleData.substr(pos + 1 + (simpleData.length() > pos + 1 && simpleData[pos + 1] == '\n'));
            }
        }
};

// ---- POCSAG ------------------------------------------------------------------------------------------------------
// Pocsag::Decoder has no MetaCollector: a finished Message hands its {address, message} map to the decoder's
// Serializer and the rendered text goes to the OUTPUT writer (reference src/pocsag_decoder/message.cpp:16-24).  The
// bank renders with the default StringSerializer on the device; for callers with another Serializer the kernel also
// leaves one event per message — {record length, number of address digits} — from which the structured map is cut
// out of the byte stream here (no separator search: message bodies may contain ';', ':' and newlines).
class PocsagReplay: public MetaReplay {
    public:
        void apply(const DecEvent*, uint32_t, std::string&) override {}
        void apply_with_output(const DecEvent* ev, uint32_t n, const uint8_t* bytes, size_t nbytes, std::string&) override {
            if (!kv_sink) return;
            size_t pos = 0;
            for (uint32_t i = 0; i < n; i++) {
                if (ev[i].kind != 1) continue;
                const size_t len = ev[i].data[0] | (size_t) ev[i].data[1] << 8;
                const size_t digits = ev[i].a;
                if (pos + len > nbytes || len < 8 + digits + 9 + 1) break;
                std::map<std::string, std::string> kv;
                kv["address"] = std::string(reinterpret_cast<const char*>(bytes + pos + 8), digits);
                kv["message"] = std::string(reinterpret_cast<const char*>(bytes + pos + 8 + digits + 9), len - (8 + digits + 9) - 1);
                emit_kv_only(kv);
                pos += len;
            }
        }
        void save(std::string&) const override {}
        bool load(const uint8_t*, size_t len) override { return len == 0; }
};

}  // namespace

MetaReplay* make_dstar_replay() { return new DstarReplay(); }
MetaReplay* make_pocsag_replay() { return new PocsagReplay(); }
MetaReplay* make_nxdn_replay() { return new NxdnReplay(); }
MetaReplay* make_dmr_replay() { return new DmrReplay(); }
MetaReplay* make_ysf_replay() { return new YsfReplay(); }

}  // namespace dh
