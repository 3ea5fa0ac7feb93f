// decoder_ops.hpp — what a protocol contributes to the generic decoder bank (decoder.cu).
#pragma once
#include "decoder.cuh"
#include "meta_replay.hpp"

namespace dh {

struct ProtoOps {
    const char* name;
    size_t state_size;                 // bytes of device state per channel
    int carry_cap;                     // symbols that may be carried between calls (multiple of 16)
    // fills `count` power-on states
    void (*init_states)(void* host_states, uint32_t count);
    // per-call capacities for a call that appends at most max_syms symbols per channel
    uint32_t (*out_bytes)(size_t max_syms);
    uint32_t (*events)(size_t max_syms);
    int (*launch)(const DecIo& io, void* d_states, const uint8_t* d_slot_filter, cudaStream_t stream);
    MetaReplay* (*make_replay)();      // may be null: protocol without metadata plane
};

const ProtoOps* dmr_ops();
const ProtoOps* pocsag_ops();
const ProtoOps* ysf_ops();
const ProtoOps* nxdn_ops();
const ProtoOps* dstar_ops();

}  // namespace dh
