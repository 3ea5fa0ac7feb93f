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

const ProtoOps* proto_ops(int proto);   // null for an unknown protocol id
const ProtoOps* dmr_ops();
const ProtoOps* pocsag_ops();
const ProtoOps* ysf_ops();
const ProtoOps* nxdn_ops();
const ProtoOps* dstar_ops();

// Device view of one of the two result sets of a decoder bank (for stages that consume the results on the device,
// e.g. the wire packing of a sharded pipe): fixed-slot rows + per-channel counts [3][channels] (out_len, ev_len, flags).
struct DecoderView {
    uint8_t* out;
    uint32_t out_cap;
    DecEvent* ev;
    uint32_t ev_cap;
    uint32_t* counts;
    uint32_t channels;
    size_t max_syms;   // per-call symbol capacity the slots were sized for
};
int decoder_view(dh_decoder* h, int set, DecoderView* view);

}  // namespace dh
