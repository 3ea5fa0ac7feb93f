// meta_replay.hpp — host-side replay of decoder event records into reference-format metadata lines.
#pragma once
#include "decoder.cuh"

#include <map>
#include <string>

namespace dh {

class MetaReplay {
    public:
        virtual ~MetaReplay() = default;
        // consumes n events of ONE channel in order and appends the resulting lines to out
        virtual void apply(const DecEvent* ev, uint32_t n, std::string& out) = 0;
        // optional second sink: the same updates as length-prefixed key/value records (for callers that run their
        // own Digiham::Serializer): per update u16 pairs, then per pair u16 klen, key, u16 vlen, value (little endian)
        std::string* kv_sink = nullptr;
    protected:
        void emit(const std::map<std::string, std::string>& kv, std::string& out);
};

MetaReplay* make_dmr_replay();
MetaReplay* make_ysf_replay();
MetaReplay* make_nxdn_replay();
MetaReplay* make_dstar_replay();

}  // namespace dh
