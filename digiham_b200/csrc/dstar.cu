// dstar.cu — D-Star decoder bank: frame sync search, radio header (K=3 Viterbi), voice frames, slow data,
// terminator; one warp per channel (sm_100a).  SURVEY.md §8(f) rank 3.
//
// Device side replaces Digiham::DStar::{SyncPhase,HeaderPhase,VoicePhase}::process (reference
// src/dstar_decoder/dstar_phase.cpp:17-149), Header::parseFromHeader incl. its 4-state Viterbi decoder
// (header.cpp:23-148), Scrambler (scrambler.cpp:6-21) and Crc::isCrcValid (crc.cpp:6-23).  Everything that is
// string handling driven by rare events — slow-data reassembly (message, header resend, DPRS / NMEA sentences,
// dstar_phase.cpp:163-278), callsign formatting (header.cpp:150-178) and DStar::MetaCollector
// (dstar_meta.cpp:5-130) — is replayed on the host from event records (meta_replay.cu).
//
// Symbols are one bit per byte (fsk_demodulator output); like the reference's byte-wise hamming_distance
// (src/lib/hamming_distance.c:3-10) a set bit 1 of a symbol byte counts as one more mismatch.
//
// The header decoder keeps the reference's exact survivor selection (ties -> predecessor k = 0, final winner =
// lowest state with the smallest metric) but stores one decision bit per state and step and traces back once,
// which yields the same bit string as the reference's register exchange.
#include "decoder_ops.hpp"

#include <cstring>

namespace dh {

constexpr int kDsCarryCap = 672;   // a header waits for more than 660 buffered symbols
constexpr int kDsSync = 24, kDsHeader = 660, kDsVoiceNeed = 120;

enum : uint8_t {
    kDsEvHeader = 1,      // a = chunk 0..3 of the 41 decoded header bytes (12 + 12 + 12 + 5); chunk 3 -> setFromHeader
    kDsEvVoiceStart = 2,  // a fresh VoicePhase: slow-data collectors start empty
    kDsEvData = 3,        // one slow-data block: mini header + 5 bytes (dstar_phase.cpp:163-211)
    kDsEvSyncDue = 4,     // a != 0: setSync("voice") first; then parseFrameData(); resetFrames()
    kDsEvReset = 5,       // MetaCollector::reset()
};

struct DstarState {
    int carry_len;
    int phase;        // 0 = SyncPhase, 1 = HeaderPhase, 2 = VoicePhase
    int frameCount;
    int syncCount;
    uint32_t first3;  // slow data of the even frame of a pair
};

#ifdef __CUDACC__
namespace {

__host__ __device__ constexpr uint32_t pack24(const int* bits) {
    uint32_t p = 0;
    for (int i = 0; i < 24; i++) p |= (uint32_t) (bits[i] & 1) << i;
    return p;
}
__host__ __device__ constexpr uint32_t header_sync_word() {
    const int b[24] = {0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0, 0};
    return pack24(b);
}
__host__ __device__ constexpr uint32_t voice_sync_word() {
    const int b[24] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0};
    return pack24(b);
}
__host__ __device__ constexpr uint32_t terminator_word(int half) {
    const int b[48] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0,
                       1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 0};
    return pack24(b + 24 * half);
}
constexpr uint32_t kHeaderSync = header_sync_word(), kVoiceSync = voice_sync_word();
constexpr uint32_t kTerm0 = terminator_word(0), kTerm1 = terminator_word(1);
constexpr uint32_t kM24 = 0xFFFFFFu;

// byte-wise hamming distance of 24 symbols (bit planes lo/hi) against a 0/1 pattern
__device__ __forceinline__ int dist24(uint32_t lo, uint32_t hi, uint32_t pattern) {
    return __popc((lo ^ pattern) & kM24) + __popc(hi & kM24);
}

// scrambler output (scrambler.cpp:10-21) for 660 bits from reset, bit i -> bit (i & 31) of word i >> 5
struct PnTable {
    uint32_t w[21];
};
__host__ __device__ constexpr PnTable make_pn() {
    PnTable t = {};
    unsigned sr = 0x7F;
    for (int i = 0; i < 660; i++) {
        const unsigned wb = (sr & 1u) ^ ((sr >> 3) & 1u);
        t.w[i >> 5] |= wb << (i & 31);
        sr = ((sr & 0x7Eu) >> 1) | (wb << 6);
    }
    return t;
}
__constant__ PnTable c_ds_pn = make_pn();

struct DCtx {
    DstarState st;
    DecWriter w;
    uint8_t* buf;    // 672 bytes of per-warp shared memory
    uint8_t* dec;    // 336 bytes: Viterbi decisions
    int lane;
};

// Header::parseFromHeader (header.cpp:23-58) on buf[0..660) (raw symbols).  On success the 41 header bytes are
// left in c.dec[0..41) and true is returned.
__device__ bool parse_header(DCtx& c) {
    const int lane = c.lane;
    uint8_t* raw = c.buf;
    // descramble in place
    for (int i = lane; i < kDsHeader; i += 32) raw[i] = (uint8_t) ((raw[i] & 1u) ^ ((c_ds_pn.w[i >> 5] >> (i & 31)) & 1u));
    __syncwarp();
    // de-interleave (header.cpp:60-74) straight into one received dibit per trellis step
    uint8_t* dib = c.dec;   // reused: dibits live in dec[0..330) until the decisions overwrite them step by step
    uint8_t mine[11];
#pragma unroll
    for (int q = 0; q < 11; q++) {
        const int p = lane + 32 * q;
        uint32_t v = 0;
        if (p < 330) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int j = 2 * p + h, i = j % 24, k = j / 24;
                v = (v << 1) | raw[i < 12 ? i * 28 + k : 12 + i * 27 + k];
            }
        }
        mine[q] = (uint8_t) v;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 11; q++) {
        const int p = lane + 32 * q;
        if (p < 330) dib[p] = mine[q];
    }
    __syncwarp();

    // 4-state Viterbi (header.cpp:78-148): state = lane & 3, every group of four lanes runs the same decode
    const int state = lane & 3;
    const uint32_t outbit = (uint32_t) (state >> 1) & 1u;
    const int p0 = (state << 1) & 2;
    auto expected = [](int prev, uint32_t ob) -> uint32_t {
        uint32_t t = ob ? 3u : 0u;
        if (prev & 1) t ^= 3u;
        if (prev & 2) t ^= 2u;
        return t;
    };
    const uint32_t e0 = expected(p0, outbit), e1 = expected(p0 | 1, outbit);
    const int base = lane & ~3;
    uint32_t metric = 0;
    for (int pos = 0; pos < 330; pos++) {
        const uint32_t in = dib[pos];
        const uint32_t m0 = __shfl_sync(0xffffffffu, metric, base | p0) + __popc(in ^ e0);
        const uint32_t m1 = __shfl_sync(0xffffffffu, metric, base | p0 | 1) + __popc(in ^ e1);
        const bool take1 = m1 < m0;
        metric = take1 ? m1 : m0;
        const uint32_t d = __ballot_sync(0xffffffffu, take1) & 0xFu;
        __syncwarp();
        if (lane == 0) c.dec[pos] = (uint8_t) d;   // dib[pos] has been consumed by every lane
    }
    __syncwarp();
    uint32_t key = (metric << 2) | (uint32_t) state;
    key = min(key, __shfl_xor_sync(0xffffffffu, key, 1));
    key = min(key, __shfl_xor_sync(0xffffffffu, key, 2));
    const uint32_t errors = key >> 2;
    // trace back (lane 0), decoded bits LSB first (header.cpp:96-100)
    uint8_t* out = c.buf;   // the raw symbols are no longer needed
    if (lane == 0) {
        for (int i = 0; i < 42; i++) out[i] = 0;
        int s = (int) (key & 3u);
        for (int pos = 329; pos >= 0; pos--) {
            out[pos >> 3] |= (uint8_t) (((s >> 1) & 1) << (pos & 7));
            s = ((s << 1) & 2) | ((c.dec[pos] >> s) & 1);
        }
    }
    __syncwarp();
    if (errors > 10) return false;
    // Crc::isCrcValid over 39 bytes against the little-endian word behind them
    uint32_t crc = 0xFFFF;
    for (int k = 0; k < 39; k++) {
        const uint32_t byte = out[k];
        for (int i = 0; i < 8; i++) {
            crc ^= (byte >> i) & 1u;
            crc = (crc & 1u) ? ((crc >> 1) ^ 0x8408u) : (crc >> 1);
        }
    }
    crc ^= 0xFFFFu;
    if (crc != ((uint32_t) out[39] | ((uint32_t) out[40] << 8))) return false;
    __syncwarp();
    for (int i = lane; i < 41; i += 32) c.dec[i] = out[i];
    __syncwarp();
    return true;
}

constexpr int kDWarps = 4;

__global__ void __launch_bounds__(kDWarps * 32) dstar_kernel(const __grid_constant__ DecIo io, DstarState* states) {
    __shared__ __align__(16) uint8_t s_buf[kDWarps][672];
    __shared__ __align__(16) uint8_t s_dec[kDWarps][336];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kDWarps + warp;
    if (ch >= io.channels) return;

    DCtx c;
    c.st = states[ch];
    c.lane = lane;
    c.buf = s_buf[warp];
    c.dec = s_dec[warp];
    c.w.out = io.out + (size_t) ch * io.out_cap;
    c.w.ev = io.ev + (size_t) ch * io.ev_cap;
    c.w.out_len = io.out_len[ch];
    c.w.ev_len = io.ev_len[ch];
    c.w.out_cap = io.out_cap;
    c.w.ev_cap = io.ev_cap;
    c.w.flags = 0;
    DstarState& s = c.st;

    uint8_t* row = io.sym + (size_t) ch * io.sym_pitch;
    const int carry_len = s.carry_len;
    const uint8_t* stream = row + (io.carry_cap - carry_len);
    const int T = carry_len + (int) min((unsigned long long) io.nsym[ch], io.sym_pitch - io.carry_cap);
    int pos = 0;

    for (;;) {
        if (s.phase == 0) {
            // SyncPhase (dstar_phase.cpp:17-35): more than 24 symbols buffered; 32 offsets per step
            const int avail = T - pos - kDsSync;
            if (avail <= 0) break;
            const int i0 = pos + lane;
            const uint8_t v0 = i0 < T ? stream[i0] : 0;
            const uint8_t v1 = i0 + 32 < T ? stream[i0 + 32] : 0;
            const uint32_t a_lo = __ballot_sync(0xffffffffu, v0 & 1), a_hi = __ballot_sync(0xffffffffu, (v0 >> 1) & 1);
            const uint32_t b_lo = __ballot_sync(0xffffffffu, v1 & 1), b_hi = __ballot_sync(0xffffffffu, (v1 >> 1) & 1);
            const uint32_t lo = __funnelshift_r(a_lo, b_lo, lane), hi = __funnelshift_r(a_hi, b_hi, lane);
            const bool is_header = dist24(lo, hi, kHeaderSync) <= 2;
            const bool is_voice = dist24(lo, hi, kVoiceSync) <= 1;
            const uint32_t hits = __ballot_sync(0xffffffffu, lane < avail && (is_header || is_voice));
            if (hits) {
                const int first = __ffs(hits) - 1;
                const bool header = __shfl_sync(0xffffffffu, (int) is_header, first) != 0;
                pos += first + kDsSync;
                if (header) {
                    s.phase = 1;
                } else {
                    // VoicePhase(0) (dstar_phase.cpp:61-63)
                    s.phase = 2;
                    s.frameCount = 0;
                    s.syncCount = 0;
                    s.first3 = 0;
                    c.w.event(lane, kDsEvVoiceStart, 0);
                }
            } else {
                pos += min(32, avail);
            }
        } else if (s.phase == 1) {
            // HeaderPhase (dstar_phase.cpp:37-59)
            if (T - pos <= kDsHeader) break;
            for (int i = lane; i < kDsHeader; i += 32) c.buf[i] = stream[pos + i];   // rare: once per transmission
            __syncwarp();
            if (!parse_header(c)) {
                pos += 1;
                s.phase = 0;
            } else {
                pos += kDsHeader;
                if (!((c.dec[0] >> 7) & 1)) {   // isVoice
                    for (int k = 0; k < 4; k++) c.w.event(lane, kDsEvHeader, 0, (uint8_t) k, 0, c.dec + 12 * k, k < 3 ? 12 : 5);
                    c.w.event(lane, kDsEvVoiceStart, 0);
                    // VoicePhase() (dstar_phase.cpp:67-70): a sync is due immediately and the header counts as one
                    s.phase = 2;
                    s.frameCount = 21;
                    s.syncCount = 1;
                    s.first3 = 0;
                } else {
                    s.phase = 0;
                }
            }
            __syncwarp();
        } else {
            // VoicePhase::process (dstar_phase.cpp:78-149): 72 voice + 24 data symbols, looks 24 symbols ahead
            if (T - pos <= kDsVoiceNeed) break;
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = lane + 32 * q;
                const uint8_t v = i < kDsVoiceNeed ? stream[pos + i] : 0;
                lo[q] = __ballot_sync(0xffffffffu, v & 1);
                hi[q] = __ballot_sync(0xffffffffu, (v >> 1) & 1);
            }
            if (s.syncCount >= 1) {
                if (c.w.out_len + 9 <= c.w.out_cap) {
                    if (lane < 9) {
                        const uint32_t word = lane < 4 ? lo[0] : (lane < 8 ? lo[1] : lo[2]);
                        c.w.out[c.w.out_len + lane] = (uint8_t) (word >> (8 * (lane & 3)));
                    }
                    c.w.out_len += 9;
                } else {
                    c.w.flags |= kFlagOutOverflow;
                }
            }
            const uint32_t d_lo = lo[2] >> 8, d_hi = hi[2] >> 8;   // symbols 72..95
            const uint32_t n_lo = lo[3], n_hi = hi[3];             // symbols 96..119
            pos += 96;
            if (dist24(d_lo, d_hi, kTerm0) + dist24(n_lo, n_hi, kTerm1) <= 1 || dist24(d_lo, d_hi, kTerm1) <= 1) {
                pos += 24;
                c.w.event(lane, kDsEvReset, 0);
                s.phase = 0;
                continue;
            }
            if (s.frameCount >= 20) {
                uint8_t set_sync = 0;
                if (dist24(d_lo, d_hi, kVoiceSync) > 1) {
                    if (--s.syncCount < 0) {
                        c.w.event(lane, kDsEvReset, 0);
                        s.phase = 0;
                        continue;
                    }
                } else {
                    if (++s.syncCount > 3) s.syncCount = 3;
                    if (s.syncCount > 1) set_sync = 1;
                }
                c.w.event(lane, kDsEvSyncDue, 0, set_sync);
                s.frameCount = 0;
            } else {
                const uint32_t bytes3 = (d_lo ^ c_ds_pn.w[0]) & kM24;   // descrambled, LSB first
                if ((s.frameCount & 1) == 0) {
                    s.first3 = bytes3;
                } else {
                    const uint32_t type = (s.first3 >> 4) & 0xFu;
                    if (type >= 3 && type <= 5) {
                        const uint8_t d[6] = {(uint8_t) s.first3, (uint8_t) (s.first3 >> 8), (uint8_t) (s.first3 >> 16),
                                              (uint8_t) bytes3, (uint8_t) (bytes3 >> 8), (uint8_t) (bytes3 >> 16)};
                        c.w.event(lane, kDsEvData, 0, 0, 0, d, 6);
                    }
                }
                s.frameCount++;
            }
        }
    }

    carry_symbols(row, io.carry_cap, carry_len, pos, T, lane);
    s.carry_len = T - pos;
    if (lane == 0) {
        states[ch] = s;
        io.out_len[ch] = c.w.out_len;
        io.ev_len[ch] = c.w.ev_len;
        if (c.w.flags) io.flags[ch] |= c.w.flags;
    }
}

}  // namespace
#endif  // __CUDACC__

namespace {

void dstar_init_states(void* host_states, uint32_t count) {
    std::memset(host_states, 0, (size_t) count * sizeof(DstarState));
}
// 9 voice bytes per 96-symbol frame
uint32_t dstar_out_bytes(size_t max_syms) { return (uint32_t) (9 * ((max_syms + kDsCarryCap) / 96 + 2)); }
// per frame at most one event; a header (684 symbols) adds five; a failed header costs 25 symbols and no event
uint32_t dstar_events(size_t max_syms) { return (uint32_t) (2 * ((max_syms + kDsCarryCap) / 96 + 2) + 16); }

int dstar_launch(const DecIo& io, void* d_states, const uint8_t*, cudaStream_t stream) {
    const unsigned grid = (io.channels + kDWarps - 1) / kDWarps;
    dstar_kernel<<<grid, kDWarps * 32, 0, stream>>>(io, static_cast<DstarState*>(d_states));
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

const ProtoOps kDstarOps = {"dstar", sizeof(DstarState), kDsCarryCap, dstar_init_states, dstar_out_bytes,
                            dstar_events, dstar_launch, make_dstar_replay};

}  // namespace

const ProtoOps* dstar_ops() { return &kDstarOps; }

}  // namespace dh
