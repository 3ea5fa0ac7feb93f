// meta_replay.hpp — host-side replay of decoder event records into reference-format metadata lines.
#pragma once
#include "decoder.cuh"

#include <string>

namespace dh {

class MetaReplay {
    public:
        virtual ~MetaReplay() = default;
        // consumes n events of ONE channel in order and appends the resulting lines to out
        virtual void apply(const DecEvent* ev, uint32_t n, std::string& out) = 0;
};

MetaReplay* make_dmr_replay();
MetaReplay* make_ysf_replay();

}  // namespace dh
