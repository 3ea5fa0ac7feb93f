// meta_replay.hpp — host-side replay of decoder event records into reference-format metadata lines.
#pragma once
#include "decoder.cuh"

#include <cstring>
#include <map>
#include <string>

namespace dh {

class MetaReplay {
    public:
        virtual ~MetaReplay() = default;
        // consumes n events of ONE channel in order and appends the resulting lines to out
        virtual void apply(const DecEvent* ev, uint32_t n, std::string& out) = 0;
        // same, for protocols whose events refer to the decoder's byte stream: `bytes` are the bytes the channel
        // produced in the same collect interval (POCSAG: record boundaries of the rendered messages)
        virtual void apply_with_output(const DecEvent* ev, uint32_t n, const uint8_t* bytes, size_t nbytes, std::string& out) {
            (void) bytes;
            (void) nbytes;
            apply(ev, n, out);
        }
        // optional second sink: the same updates as length-prefixed key/value records (for callers that run their
        // own Digiham::Serializer): per update u16 pairs, then per pair u16 klen, key, u16 vlen, value (little endian)
        std::string* kv_sink = nullptr;
        // collector state (what the reference keeps in its MetaCollector / Slot / TalkerAliasCollector objects) as a
        // flat byte string, for dh_decoder_state_export / _import; load returns false on a malformed blob
        virtual void save(std::string& blob) const = 0;
        virtual bool load(const uint8_t* data, size_t len) = 0;
    protected:
        void emit(const std::map<std::string, std::string>& kv, std::string& out);
        void emit_kv_only(const std::map<std::string, std::string>& kv);   // key/value record only, no text line
};

// Field-by-field (de)serialisation used by the save / load of the collectors: every class lists its members once in
// a `fields(archive)` template, StateWriter appends them, StateReader reads them back in the same order.
struct StateWriter {
    std::string& out;
    void pod(const void* p, size_t n) { out.append(static_cast<const char*>(p), n); }
    template <class T> void val(const T& v) { pod(&v, sizeof(T)); }
    void str(const std::string& s) {
        const uint32_t n = (uint32_t) s.size();
        val(n);
        out.append(s);
    }
};
struct StateReader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    void pod(void* dst, size_t n) {
        if (!ok || (size_t) (end - p) < n) {
            ok = false;
            return;
        }
        std::memcpy(dst, p, n);
        p += n;
    }
    template <class T> void val(T& v) { pod(&v, sizeof(T)); }
    void str(std::string& s) {
        uint32_t n = 0;
        val(n);
        if (!ok || (size_t) (end - p) < n) {
            ok = false;
            return;
        }
        s.assign(reinterpret_cast<const char*>(p), n);
        p += n;
    }
};

// One metadata update written straight into the channel's text (`key:value;...\n`, reference src/lib/meta.cpp:8-17) and,
// when asked for, into the key/value record stream — without the std::map the reference builds per update (the
// replay of thousands of channels per step is host time that competes with the uploads of the next step).  The
// caller adds the pairs in std::map order, i.e. sorted by key, which is the order the StringSerializer emits.
class MetaLine {
    public:
        MetaLine(std::string& text, std::string* kv): text(text), kv(kv) {
            if (kv) {
                count_pos = kv->size();
                kv->append(2, '\0');
            }
        }
        void add(const char* key, const char* val, size_t vlen) {
            const size_t klen = std::strlen(key);
            if (pairs++) text.push_back(';');
            text.append(key, klen);
            text.push_back(':');
            text.append(val, vlen);
            if (kv) {
                put16(klen);
                kv->append(key, klen);
                put16(vlen);
                kv->append(val, vlen);
            }
        }
        void add(const char* key, const std::string& val) { add(key, val.data(), val.size()); }
        void add(const char* key, const char* val) { add(key, val, std::strlen(val)); }
        void add_uint(const char* key, unsigned long long v) {
            char buf[24];
            int n = 0;
            do {
                buf[n++] = (char) ('0' + v % 10);
                v /= 10;
            } while (v);
            char out[24];
            for (int i = 0; i < n; i++) out[i] = buf[n - 1 - i];
            add(key, out, (size_t) n);
        }
        void finish() {
            text.push_back('\n');
            if (kv) {
                (*kv)[count_pos] = (char) (pairs & 0xFF);
                (*kv)[count_pos + 1] = (char) ((pairs >> 8) & 0xFF);
            }
        }
    private:
        void put16(size_t v) {
            kv->push_back((char) (v & 0xFF));
            kv->push_back((char) ((v >> 8) & 0xFF));
        }
        std::string& text;
        std::string* kv;
        size_t count_pos = 0;
        unsigned pairs = 0;
};

MetaReplay* make_dmr_replay();
MetaReplay* make_ysf_replay();
MetaReplay* make_nxdn_replay();
MetaReplay* make_dstar_replay();
MetaReplay* make_pocsag_replay();

}  // namespace dh
