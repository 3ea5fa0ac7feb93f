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

MetaReplay* make_dmr_replay();
MetaReplay* make_ysf_replay();
MetaReplay* make_nxdn_replay();
MetaReplay* make_dstar_replay();
MetaReplay* make_pocsag_replay();

}  // namespace dh
