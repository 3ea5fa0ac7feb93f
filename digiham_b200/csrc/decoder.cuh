// decoder.cuh — plumbing shared by the protocol decoder banks (DMR / YSF / POCSAG), sm_100a.
//
// A decoder bank replaces N instances of Digiham::Decoder (reference include/decoder.hpp:17-30,
// src/lib/decoder.cpp:21-32): a phase state machine per channel that consumes demodulated symbols (one byte
// each) and emits (a) a byte stream and (b) metadata updates.  On the GPU one warp owns one channel:
//   * symbols live in per-channel rows with the unconsumed tail carried right-aligned in front of the column
//     where the producer (K2) writes the next chunk, so the logical stream is contiguous;
//   * output bytes are appended to per-channel rows, metadata leaves the GPU as compact 16-byte event records
//     which the host replays into the reference's `key:value;...\n` lines (src/lib/meta.cpp:8-17).
#pragma once
#include "common.cuh"

#include <string>
#include <vector>

namespace dh {

struct DecEvent {
    uint8_t kind;
    uint8_t slot;
    uint8_t a;
    uint8_t b;
    uint8_t data[12];
};
static_assert(sizeof(DecEvent) == 16, "event records are 16 bytes");

constexpr uint32_t kFlagOutOverflow = 1u;
constexpr uint32_t kFlagEventOverflow = 2u;

struct DecIo {
    uint8_t* sym;                 // [channels][sym_pitch]; new symbols start at column carry_cap
    unsigned long long sym_pitch;
    const uint32_t* nsym;         // [channels] symbols appended by the producer for this call
    uint8_t* out;                 // [channels][out_cap]
    uint32_t* out_len;            // [channels] bytes appended since the last collect
    DecEvent* ev;                 // [channels][ev_cap]
    uint32_t* ev_len;             // [channels]
    uint32_t* flags;              // [channels] kFlag* bits
    uint32_t out_cap;
    uint32_t ev_cap;
    int carry_cap;
    int channels;
};

#ifdef __CUDACC__

// Appends bytes / events of one channel; every lane of the owning warp holds the same counters.
struct DecWriter {
    uint8_t* out;
    DecEvent* ev;
    uint32_t out_len, ev_len, out_cap, ev_cap, flags;

    __device__ __forceinline__ void event(int lane, uint8_t kind, uint8_t slot, uint8_t a = 0, uint8_t b = 0,
                                          const uint8_t* data = nullptr, int ndata = 0) {
        if (ev_len >= ev_cap) {
            flags |= kFlagEventOverflow;
            return;
        }
        if (lane == 0) {
            DecEvent e;
            e.kind = kind;
            e.slot = slot;
            e.a = a;
            e.b = b;
#pragma unroll
            for (int i = 0; i < 12; i++) e.data[i] = (data != nullptr && i < ndata) ? data[i] : 0;
            ev[ev_len] = e;
        }
        ev_len++;
    }
};

// move the unconsumed tail [pos, T) of a symbol row right-aligned in front of column carry_cap (dst <= src)
__device__ __forceinline__ void carry_symbols(uint8_t* row, int carry_cap, int carry_len, int pos, int T, int lane) {
    const int keep = T - pos;
    const uint8_t* src = row + (carry_cap - carry_len) + pos;
    uint8_t* dst = row + carry_cap - keep;
    if (dst == src) return;
    for (int o = 0; o < keep; o += 32) {
        const int idx = o + lane;
        const uint8_t v = idx < keep ? src[idx] : 0;
        __syncwarp();
        if (idx < keep) dst[idx] = v;
        __syncwarp();
    }
}

__device__ __forceinline__ uint32_t parity32(uint32_t v) { return __popc(v) & 1u; }

// Stage `count` symbols starting at src into a 4-byte aligned shared buffer with aligned 32-bit loads; returns the
// address of the first symbol inside the buffer (buffer size >= count + 4; symbol rows are 16-byte aligned, so no
// word leaves the row).  The caller synchronises the warp afterwards.
__device__ __forceinline__ const uint8_t* stage_symbols(uint8_t* buf, const uint8_t* src, int count, int lane) {
    const int mis = (int) (reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t* src_w = reinterpret_cast<const uint32_t*>(src - mis);
    uint32_t* dst_w = reinterpret_cast<uint32_t*>(buf);
    const int nwords = (count + mis + 3) >> 2;
    for (int i = lane; i < nwords; i += 32) dst_w[i] = src_w[i];
    return buf + mis;
}

#endif  // __CUDACC__

// Host side of a bank: per-channel accumulated results between dh_decoder_collect calls.
struct ChannelResult {
    std::string bytes;
    std::string meta;
    std::string meta_kv;   // same updates as key/value records (see MetaReplay::kv_sink)
};

class MetaReplay;
class WorkerPool;   // result_sink.cu: persistent host threads for the per-channel replay

// Where device result blocks end up on the host: per-channel byte streams + the metadata lines replayed from the
// 16-byte event records.  One sink serves a decoder bank (its own channels) or the gathering rank of a sharded
// pipe (the channels of every rank, shard.cu).
struct ResultSink {
    uint32_t channels = 0;
    std::vector<ChannelResult> results;
    std::vector<MetaReplay*> replay;      // null entries: protocol without metadata plane
    uint32_t* h_counts = nullptr;         // pinned, [3][channels]
    uint8_t* h_out = nullptr;             // pinned staging, grown on demand
    size_t h_out_bytes = 0;
    DecEvent* h_ev = nullptr;
    size_t h_ev_bytes = 0;
    WorkerPool* pool = nullptr;           // created on first use, joined by release()
    uint8_t* d_compact = nullptr;         // device staging: the used widths of all rows, densely packed, so that the
    size_t d_compact_bytes = 0;           // read-back is ONE contiguous copy instead of thousands of short rows
    uint64_t total_bytes = 0, total_meta = 0, total_events = 0, total_d2h = 0;
    // key/value records beside the text lines (dh_decoder_meta_kv): kept for small banks (the facade's one-channel
    // banks apply their own Serializer to them), off for large ones unless asked for (dh_decoder_set_meta_kv)
    bool want_kv = false;

    int init(int proto, uint32_t nchannels);
    // Reads one device result block — counts [3][n] (out_len, ev_len, flags), byte rows [n][out_pitch], event rows
    // [n][ev_pitch records] — of the sink channels [c0, c0 + n): copies the used widths to pinned memory on `st`
    // (synchronises it), appends the bytes and replays the events in order.  The flag words are OR-ed into *flags.
    int ingest(const uint32_t* d_counts, const uint8_t* d_out, size_t out_pitch, const DecEvent* d_ev, size_t ev_pitch,
               uint32_t n, uint32_t c0, cudaStream_t st, uint32_t* flags);
    // Several blocks at once (the gathering rank of a sharded pipe: one block per rank): all counts are fetched
    // with one synchronisation, all rows with a second one, then every channel is replayed in one parallel pass.
    struct Block {
        const uint32_t* d_counts;
        const uint8_t* d_out;
        const DecEvent* d_ev;
        uint32_t n;    // channels of the block
        uint32_t c0;   // first sink channel
    };
    int ingest_blocks(const Block* blocks, int nblocks, size_t out_pitch, size_t ev_pitch, cudaStream_t st, uint32_t* flags);
    void clear();
    void release();
};

}  // namespace dh
