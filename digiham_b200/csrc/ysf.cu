// ysf.cu — K3/K5 for Yaesu System Fusion: sync search, FICH, voice/data channel extraction; one warp per channel
// (sm_100a).
//
// Device side replaces Digiham::Ysf::{SyncPhase,FramePhase}::process, Fich::parse, decode_trellis, golay_24_12,
// crc16 and decode_whitening (reference src/ysf_decoder/ysf_phase.cpp:16-349, fich.cpp:12-66, trellis.c:8-109,
// golay_24_12.c, crc16.c:3-22, whitening.c:6-22).  Callsign strings, the DT1/DT2 data-frame collector and GPS
// parsing (ysf_phase.cpp:351-361, data.cpp:15-88, gps.cpp:7-105, ysf_meta.cpp) are string/float handling driven
// by rare events and are replayed on the host (meta_replay.cu).
//
// K5, the rate-1/2 K=5 hard-decision Viterbi decoder, runs one trellis state per lane (16 states; both half-warps
// carry the same decode): per step each lane fetches the two predecessor metrics with __shfl_sync, compare-selects
// (ties -> predecessor k = 0, exactly like the reference's strict `<`), then pulls the selected survivor —
// register exchange, the whole decoded bit string travels with the state — word by word through __shfl_sync.
// Metrics are kept modulo 256 like the reference's uint8_t (trellis.c:28,68).
#include "decoder_ops.hpp"
#include "test_hooks.hpp"
#include "viterbi.cuh"
#include "crc_par.cuh"

#define DH_TABLES_NO_HOST_ARRAYS
#include "tables.inc"

#include <cstring>

namespace dh {

constexpr int kYsfCarryCap = 496;
constexpr int kYsfFrame = 480, kYsfSync = 20, kYsfFich = 100;

enum : uint8_t {
    kYsfEvMode = 1,        // setMode(a): 1 "V1", 2 "DN", 3 "VW", 4 "FR data"
    kYsfEvReset = 2,       // MetaCollector::reset()
    kYsfEvHold = 3,        // hold()
    kYsfEvField = 4,       // a: 0 destination, 1 source, 2 down, 3 up; data = 10 raw bytes (treatYsfString on host)
    kYsfEvRelease = 5,     // release()
    kYsfEvDcReset = 6,     // dataCollector->reset()
    kYsfEvDcCollect = 7,   // dataCollector->collect(data, a)
    kYsfEvDcCheck = 8,     // if (hasCollected(2)) { getDataFrame -> setGps }
};

struct YsfState {
    int carry_len;
    int phase;            // 0 = SyncPhase, 1 = FramePhase
    int syncCount;
    int has_fich;
    uint32_t fich;        // runningFich
    int expectSubFrame;
    int dc_next;          // mirror of DataCollector::nextOffset
    int m_mode;           // mirror of the collector's mode (0 = empty)
    int m_any;            // some field / coordinate may be set on the host
};

#ifdef __CUDACC__
namespace {

__constant__ uint32_t c_golay24_lut[4096] = DH_GOLAY_24_12_LUT_INIT;
__constant__ uint32_t c_golay24_h[12] = DH_GOLAY_24_12_H_INIT;

// sync word D471C9634D as dibit planes, symbol i -> bit i (ysf_phase.hpp:21)
__host__ __device__ constexpr uint32_t sync_plane(unsigned long long hex40, int which) {
    uint32_t p = 0;
    for (int i = 0; i < 20; i++) {
        const unsigned dibit = (unsigned) ((hex40 >> (38 - 2 * i)) & 3ull);
        p |= ((which ? (dibit >> 1) : dibit) & 1u) << i;
    }
    return p;
}
constexpr uint32_t kSyncHi = sync_plane(0xD471C9634Dull, 1);
constexpr uint32_t kSyncLo = sync_plane(0xD471C9634Dull, 0);

__device__ __forceinline__ bool is_sync(uint32_t hi, uint32_t lo) {
    const uint32_t m = 0xFFFFFu;
    return __popc((hi ^ kSyncHi) & m) + __popc((lo ^ kSyncLo) & m) <= 3;
}

// PN9 whitening sequence (whitening.c:7-20): register 0b111001001, output bit 0, feedback bit0 ^ bit4 into bit 8.
// pn_bit(i) for i < 160, packed MSB-first per 32-bit word like the data it is XORed with.
struct PnTable {
    uint32_t w[5];
};
__host__ __device__ constexpr PnTable make_pn() {
    PnTable t = {{0, 0, 0, 0, 0}};
    unsigned wsr = 0x1C9;
    for (int i = 0; i < 160; i++) {
        const unsigned wb = wsr & 1u;
        t.w[i >> 5] |= wb << (31 - (i & 31));
        const unsigned fb = ((wsr >> 4) & 1u) ^ wb;
        wsr = ((wsr & 0x1FEu) >> 1) | (fb << 8);
    }
    return t;
}
__constant__ PnTable c_pn = make_pn();

// inverse of v2_voice_mapping (ysf_phase.hpp:45-51): output bit p carries voice bit c_v2_inverse[p]
__constant__ uint8_t c_v2_inverse[49] = {0,  18, 36, 1,  19, 37, 2,  20, 38, 3,  21, 39, 4,  22, 40, 5,  23,
                                        41, 6,  24, 42, 7,  25, 43, 8,  26, 44, 9,  27, 45, 10, 28, 46, 11,
                                        29, 47, 12, 30, 48, 13, 31, 14, 32, 15, 33, 16, 34, 17, 35};

// crc16_checksum (crc16.c:3-19) over 32 / 80 / 160 message bits, evaluated warp-parallel (crc_par.cuh)
__constant__ CrcTable<32> c_crc32 = make_crc_table<0, 32>();
__constant__ CrcTable<80> c_crc80 = make_crc_table<0, 80>();
__constant__ CrcTable<160> c_crc160 = make_crc_table<0, 160>();
struct YsfCrcTables {   // copy in shared memory: lanes index the tables with different offsets
    uint16_t t32[32], t80[80], t160[160];
};

__device__ __forceinline__ bool fec_golay24(uint32_t& w) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) s = (s << 1) | parity32(c_golay24_h[k] & w);
    if (s == 0) return true;
    const uint32_t e = c_golay24_lut[s];
    w ^= e;
    return e != 0;
}

struct YCtx {
    YsfState st;
    DecWriter w;
    const YsfCrcTables* crc;   // shared-memory copy of the CRC tables
    const uint8_t* fr;   // 480 staged dibits of the frame (shared memory)
    uint8_t* scratch;    // 384 bytes of per-warp scratch (shared memory): two dibit arrays for the paired Viterbi
    int lane;
};

__device__ __forceinline__ void meta_mode(YCtx& c, int mode) {
    if (c.st.m_mode != mode) {
        c.w.event(c.lane, kYsfEvMode, 0, (uint8_t) mode);
        c.st.m_mode = mode;
    }
}
__device__ __forceinline__ void meta_reset(YCtx& c) {
    if (c.st.m_mode != 0 || c.st.m_any) {
        c.w.event(c.lane, kYsfEvReset, 0);
        c.st.m_mode = 0;
        c.st.m_any = 0;
    }
}
__device__ __forceinline__ void meta_field(YCtx& c, int which, const uint8_t* bytes10) {
    c.w.event(c.lane, kYsfEvField, 0, (uint8_t) which, 0, bytes10, 10);
    c.st.m_any = 1;
}

// words (MSB-first bit string) -> bytes
__device__ __forceinline__ uint8_t word_byte(const uint32_t* words, int k) {
    return (uint8_t) (words[k >> 2] >> (24 - 8 * (k & 3)));
}

// Fich::parse (fich.cpp:12-52) behind the Viterbi decoder: `words` = the 100 decoded bits
__device__ bool parse_fich(YCtx& c, const uint32_t* words, uint32_t& fich_out) {
    uint32_t g[4];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        g[i] = ((uint32_t) word_byte(words, 3 * i) << 16) | ((uint32_t) word_byte(words, 3 * i + 1) << 8) |
               word_byte(words, 3 * i + 2);
        ok &= fec_golay24(g[i]);
    }
    if (!ok) return false;
    const uint32_t fich = ((g[0] & 0xFFF000u) << 8) | ((g[1] & 0xFFF000u) >> 4) | ((g[2] & 0xFF0000u) >> 16);
    const uint32_t checksum = (g[2] & 0x00F000u) | ((g[3] & 0xFFF000u) >> 12);
    // crc16 over the four FICH bytes, big endian (fich.cpp:44-49)
    if (crc_parallel<32>(&fich, c.crc->t32, c_crc32.c, c.lane) != checksum) return false;
    fich_out = fich;
    return true;
}

// FramePhase::decodeV2VoicePayload (ysf_phase.cpp:180-256): 52 dibits -> 7 bytes
__device__ void v2_voice(YCtx& c, const uint8_t* in, uint8_t* out8) {
    uint8_t* bits = c.scratch;          // 104 de-interleaved, de-whitened bits
    uint8_t* vbit = c.scratch + 112;    // 49 voice bits
    for (int k = c.lane; k < 104; k += 32) {
        const int o = (k * 4) % 104 + (k * 4) / 104;
        const uint8_t d = in[o >> 1];
        uint32_t b = (o & 1) ? (d & 1u) : ((d >> 1) & 1u);
        b ^= (c_pn.w[k >> 5] >> (31 - (k & 31))) & 1u;
        bits[k] = (uint8_t) b;
    }
    __syncwarp();
    for (int i = c.lane; i < 49; i += 32) {
        if (i < 27) {
            const int t = bits[3 * i] + bits[3 * i + 1] + bits[3 * i + 2];
            vbit[i] = t >= 2;           // tribit majority (ysf_phase.hpp:43)
        } else {
            vbit[i] = bits[i + 54];     // bits 81..102 pass through
        }
    }
    __syncwarp();
    if (c.lane < 7) {
        uint32_t b = 0;
        for (int q = 0; q < 8; q++) {
            const int p = c.lane * 8 + q;
            b = (b << 1) | (p < 49 ? vbit[c_v2_inverse[p]] : 0u);
        }
        out8[1 + c.lane] = (uint8_t) b;
    }
    __syncwarp();
}

// de-whiten the first nbits of an MSB-first bit string held in words, return byte k
__device__ __forceinline__ uint8_t dewhitened_byte(const uint32_t* words, int k) {
    return (uint8_t) ((words[k >> 2] ^ c_pn.w[k >> 2]) >> (24 - 8 * (k & 3)));
}

// FramePhase::decodeV2DataChannel (ysf_phase.cpp:258-306) behind the Viterbi decoder: `words` = its 100 decoded bits
__device__ void v2_data_channel(YCtx& c, const uint32_t* words, int frameNumber) {
    uint8_t raw[12];
#pragma unroll
    for (int k = 0; k < 12; k++) raw[k] = word_byte(words, k);
    const uint32_t checksum = ((uint32_t) raw[10] << 8) | raw[11];
    if (crc_parallel<80>(words, c.crc->t80, c_crc80.c, c.lane) != checksum) return;
    // decode_whitening(..., 100): the first 100 bits are de-whitened, i.e. all of the 10 bytes used below
    uint8_t dch[10];
#pragma unroll
    for (int k = 0; k < 10; k++) dch[k] = dewhitened_byte(words, k);
    if (frameNumber < 6) {
        if (frameNumber <= 3) meta_field(c, frameNumber, dch);
        if (c.st.dc_next != 0) {
            c.w.event(c.lane, kYsfEvDcReset, 0);
            c.st.dc_next = 0;
        }
    }
    if (frameNumber >= 6 && frameNumber < 8) {
        const int offset = frameNumber - 6;
        c.w.event(c.lane, kYsfEvDcCollect, 0, (uint8_t) offset, 0, dch, 10);
        c.st.dc_next = (offset != c.st.dc_next) ? 0 : offset + 1;   // DataCollector::collect (data.cpp:55-67)
    }
    if (c.st.dc_next >= 2) {
        c.w.event(c.lane, kYsfEvDcCheck, 0);
        c.st.m_any = 1;
    }
}

// FramePhase::decodeHeaderDataChannel (ysf_phase.cpp:317-349) behind the Viterbi decoder: `words` = 180 decoded bits
__device__ bool header_data_channel(YCtx& c, const uint32_t* words, uint8_t* dch20) {
    uint8_t raw[22];
#pragma unroll
    for (int k = 0; k < 22; k++) raw[k] = word_byte(words, k);
    const uint32_t checksum = ((uint32_t) raw[20] << 8) | raw[21];
    if (crc_parallel<160>(words, c.crc->t160, c_crc160.c, c.lane) != checksum) return false;
#pragma unroll
    for (int k = 0; k < 20; k++) dch20[k] = dewhitened_byte(words, k);
    return true;
}

__device__ __forceinline__ bool out_room(YCtx& c, uint32_t n) {
    if (c.w.out_len + n <= c.w.out_cap) return true;
    c.w.flags |= kFlagOutOverflow;
    return false;
}

// FramePhase::process (ysf_phase.cpp:45-172).  Returns false when the phase fell back to SyncPhase.
__device__ bool ysf_frame(YCtx& c) {
    YsfState& s = c.st;
    const int lane = c.lane;
    uint32_t hi, lo;
    {
        const uint8_t v = lane < kYsfSync ? c.fr[lane] : 0;
        hi = __ballot_sync(0xffffffffu, (v >> 1) & 1);
        lo = __ballot_sync(0xffffffffu, v & 1);
    }
    if (is_sync(hi, lo)) {
        if (++s.syncCount > 12) s.syncCount = 12;
    } else if (--s.syncCount < 0) {
        meta_reset(c);
        return false;
    }

    // K5: the FICH and — speculatively, the second half-warp would otherwise repeat the first — the data channel
    // of a V/D2 frame are decoded side by side (fich.cpp:12-22, ysf_phase.cpp:258-272)
    const uint8_t* payload = c.fr + kYsfSync + kYsfFich;
    uint32_t fich_words[4], dch_words[4], vm_a, vm_b;
    {
        const uint8_t* data = c.fr + kYsfSync;
        uint8_t* dib = c.scratch;
        for (int i = lane; i < 100; i += 32) {
            dib[i] = data[(i * 20) % 100 + (i * 20) / 100] & 3u;
            dib[192 + i] = payload[(i % 5) * 72 + (i * 2) / 10] & 3u;
        }
        __syncwarp();
        viterbi_pair<100, false>(dib, dib + 192, lane, fich_words, dch_words, vm_a, vm_b);
        __syncwarp();
    }
    uint32_t fich = 0;
    const bool fich_ok = parse_fich(c, fich_words, fich);
    if (fich_ok) {
        s.has_fich = 1;
        s.fich = fich;
    }
    if (s.has_fich) {
        const int frameType = (s.fich >> 30) & 3;
        const int dataType = (s.fich >> 8) & 3;
        if (frameType == 1) {
            if (dataType == 0) {          // V/D mode type 1
                meta_mode(c, 1);
                if (out_room(c, 50)) {
                    uint8_t* o = c.w.out + c.w.out_len;
                    // 5 x (mode byte + 9 bytes); each byte only keeps its last dibit (sic, ysf_phase.cpp:176)
                    for (int e = lane; e < 50; e += 32) {
                        const int blk = e / 10, j = e % 10;
                        o[e] = j == 0 ? (uint8_t) dataType : (uint8_t) (payload[36 + blk * 72 + 4 * (j - 1) + 3] & 3u);
                    }
                    c.w.out_len += 50;
                }
            } else if (dataType == 2) {   // V/D mode type 2
                meta_mode(c, 2);
                if (out_room(c, 40)) {
                    uint8_t* o = c.w.out + c.w.out_len;
                    for (int i = 0; i < 5; i++) {
                        if (lane == 0) o[i * 8] = (uint8_t) dataType;
                        v2_voice(c, payload + 20 + i * 72, o + i * 8);
                    }
                    c.w.out_len += 40;
                }
                if (fich_ok) v2_data_channel(c, dch_words, (int) ((fich >> 19) & 7));
            } else if (dataType == 3) {   // voice full rate
                meta_mode(c, 3);
                const int start = s.expectSubFrame ? 3 : 0;
                s.expectSubFrame = 0;
                const uint32_t nbytes = (uint32_t) (5 - start) * 19u;
                if (out_room(c, nbytes)) {
                    uint8_t* o = c.w.out + c.w.out_len;
                    for (int e = lane; e < (int) nbytes; e += 32) {
                        const int blk = start + e / 19, j = e % 19;
                        if (j == 0) {
                            o[e] = (uint8_t) dataType;
                        } else {
                            const uint8_t* p = payload + blk * 72 + 4 * (j - 1);
                            o[e] = (uint8_t) (((p[0] & 3u) << 6) | ((p[1] & 3u) << 4) | ((p[2] & 3u) << 2) | (p[3] & 3u));
                        }
                    }
                    c.w.out_len += nbytes;
                }
            } else {                      // data full rate: not decoded
                meta_mode(c, 4);
            }
        } else if (frameType == 0) {      // header
            meta_reset(c);
            c.w.event(lane, kYsfEvHold, 0);
            // both 180-step data channels of the header in one paired decode (ysf_phase.cpp:139-161, 317-331)
            uint32_t h0[6], h1[6];
            {
                uint8_t* dib = c.scratch;
                for (int i = lane; i < 180; i += 32) {
                    const int streampos = (i % 9) * 20 + i / 9;
                    const int at = (streampos / 36) * 72 + streampos % 36;
                    dib[i] = payload[at] & 3u;
                    dib[192 + i] = payload[36 + at] & 3u;
                }
                __syncwarp();
                viterbi_pair<180, false>(dib, dib + 192, lane, h0, h1, vm_a, vm_b);
                __syncwarp();
            }
            uint8_t dch[20];
            if (header_data_channel(c, h0, dch)) {
                meta_field(c, 0, dch);
                meta_field(c, 1, dch + 10);
            }
            if (header_data_channel(c, h1, dch)) {
                meta_field(c, 2, dch);
                meta_field(c, 3, dch + 10);
            }
            c.w.event(lane, kYsfEvRelease, 0);
            s.expectSubFrame = 1;
        } else if (frameType == 2) {      // terminator
            meta_reset(c);
        }
    }
    return true;
}

constexpr int kYWarps = 4;

__global__ void __launch_bounds__(kYWarps * 32) ysf_kernel(const __grid_constant__ DecIo io, YsfState* states) {
    __shared__ __align__(16) uint8_t s_fr[kYWarps][kYsfFrame + 8];
    __shared__ __align__(16) uint8_t s_scratch[kYWarps][384];
    __shared__ YsfCrcTables s_crc;
    for (int i = threadIdx.x; i < 160; i += kYWarps * 32) {
        if (i < 32) s_crc.t32[i] = c_crc32.t[i];
        if (i < 80) s_crc.t80[i] = c_crc80.t[i];
        s_crc.t160[i] = c_crc160.t[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kYWarps + warp;
    if (ch >= io.channels) return;

    YCtx c;
    c.st = states[ch];
    c.lane = lane;
    c.crc = &s_crc;
    c.fr = s_fr[warp];
    c.scratch = s_scratch[warp];
    c.w.out = io.out + (size_t) ch * io.out_cap;
    c.w.ev = io.ev + (size_t) ch * io.ev_cap;
    c.w.out_len = io.out_len[ch];
    c.w.ev_len = io.ev_len[ch];
    c.w.out_cap = io.out_cap;
    c.w.ev_cap = io.ev_cap;
    c.w.flags = 0;

    uint8_t* row = io.sym + (size_t) ch * io.sym_pitch;
    const int carry_len = c.st.carry_len;
    const uint8_t* stream = row + (io.carry_cap - carry_len);
    const int T = carry_len + (int) min((unsigned long long) io.nsym[ch], io.sym_pitch - io.carry_cap);
    int pos = 0;

    for (;;) {
        if (c.st.phase == 0) {
            // SyncPhase (ysf_phase.cpp:20-34): more than 20 symbols buffered, sync at the read pointer
            const int avail = T - pos - kYsfSync;
            if (avail <= 0) break;
            const int i0 = pos + lane;
            const uint8_t v0 = i0 < T ? stream[i0] : 0;
            const uint8_t v1 = i0 + 32 < T ? stream[i0 + 32] : 0;
            const uint32_t a_hi = __ballot_sync(0xffffffffu, (v0 >> 1) & 1);
            const uint32_t a_lo = __ballot_sync(0xffffffffu, v0 & 1);
            const uint32_t b_hi = __ballot_sync(0xffffffffu, (v1 >> 1) & 1);
            const uint32_t b_lo = __ballot_sync(0xffffffffu, v1 & 1);
            const bool hit = lane < avail && is_sync(__funnelshift_r(a_hi, b_hi, lane), __funnelshift_r(a_lo, b_lo, lane));
            const uint32_t hits = __ballot_sync(0xffffffffu, hit);
            if (hits) {
                pos += __ffs(hits) - 1;
                // fresh FramePhase (ysf_phase.hpp:53-56)
                c.st.phase = 1;
                c.st.syncCount = 0;
                c.st.has_fich = 0;
                c.st.expectSubFrame = 0;
                if (c.st.dc_next != 0) {
                    c.w.event(lane, kYsfEvDcReset, 0);
                    c.st.dc_next = 0;
                }
            } else {
                pos += min(32, avail);
            }
        } else {
            if (T - pos <= kYsfFrame) break;
            c.fr = stage_symbols(s_fr[warp], stream + pos, kYsfFrame, lane);
            __syncwarp();
            if (ysf_frame(c)) pos += kYsfFrame;
            else c.st.phase = 0;
            __syncwarp();
        }
    }

    carry_symbols(row, io.carry_cap, carry_len, pos, T, lane);
    c.st.carry_len = T - pos;
    if (lane == 0) {
        states[ch] = c.st;
        io.out_len[ch] = c.w.out_len;
        io.ev_len[ch] = c.w.ev_len;
        if (c.w.flags) io.flags[ch] |= c.w.flags;
    }
}

}  // namespace
#endif  // __CUDACC__

namespace {

void ysf_init_states(void* host_states, uint32_t count) {
    std::memset(host_states, 0, (size_t) count * sizeof(YsfState));
}
// at most 95 bytes (voice full rate) per 480-symbol frame
uint32_t ysf_out_bytes(size_t max_syms) { return (uint32_t) (95 * ((max_syms + kYsfCarryCap) / kYsfFrame + 2)); }
// per frame: mode, reset, hold, 4 fields, release, collector reset/collect/check
uint32_t ysf_events(size_t max_syms) { return (uint32_t) (12 * ((max_syms + kYsfCarryCap) / kYsfFrame + 2) + 8); }

int ysf_launch(const DecIo& io, void* d_states, const uint8_t*, cudaStream_t stream) {
    const unsigned grid = (io.channels + kYWarps - 1) / kYWarps;
    ysf_kernel<<<grid, kYWarps * 32, 0, stream>>>(io, static_cast<YsfState*>(d_states));
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

const ProtoOps kYsfOps = {"ysf", sizeof(YsfState), kYsfCarryCap, ysf_init_states, ysf_out_bytes, ysf_events,
                          ysf_launch, make_ysf_replay};

}  // namespace

const ProtoOps* ysf_ops() { return &kYsfOps; }

// ---- device-level test hooks (dh_test_fec code 6, dh_test_viterbi variants 0 / 1) --------------------------------
namespace {

__global__ void ysf_test_golay24_kernel(uint32_t* words, uint8_t* ok, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w = words[i];
    const bool r = fec_golay24(w);
    words[i] = w;
    ok[i] = r ? 1 : 0;
}

// one warp decodes inputs 2w and 2w + 1 side by side, exactly like ysf_frame does (viterbi_pair)
template <int STEPS>
__global__ void ysf_test_viterbi_kernel(const uint8_t* dibits, uint32_t n, uint32_t* words, uint32_t* metric) {
    constexpr int NW = (STEPS + 31) / 32;
    __shared__ uint8_t stage[4][2][192];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t a = (blockIdx.x * 4 + wib) * 2, b = a + 1;
    if (a >= n) return;
    const uint32_t bb = b < n ? b : a;
    for (int i = lane; i < STEPS; i += 32) {
        stage[wib][0][i] = dibits[(size_t) a * STEPS + i];
        stage[wib][1][i] = dibits[(size_t) bb * STEPS + i];
    }
    __syncwarp();
    uint32_t wa[NW], wb[NW], ma, mb;
    viterbi_pair<STEPS, false>(stage[wib][0], stage[wib][1], lane, wa, wb, ma, mb);
    if (lane == 0) {
        for (int k = 0; k < NW; k++) {
            words[(size_t) a * NW + k] = wa[k];
            if (b < n) words[(size_t) b * NW + k] = wb[k];
        }
        metric[a] = ma;
        if (b < n) metric[b] = mb;
    }
}

}  // namespace

namespace test {

int ysf_golay24(uint32_t* d_words, uint8_t* d_ok, uint32_t n, cudaStream_t st) {
    if (n == 0) return DH_OK;
    ysf_test_golay24_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_words, d_ok, n);
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

int ysf_viterbi(int steps, const uint8_t* d_dibits, uint32_t n, uint32_t* d_words, uint32_t* d_metric, cudaStream_t st) {
    DH_REQUIRE(steps == 100 || steps == 180, DH_E_INVALID, "dh_test_viterbi: YSF decodes 100 or 180 steps");
    if (n == 0) return DH_OK;
    const unsigned grid = ((n + 1) / 2 + 3) / 4;
    if (steps == 100) ysf_test_viterbi_kernel<100><<<grid, 128, 0, st>>>(d_dibits, n, d_words, d_metric);
    else ysf_test_viterbi_kernel<180><<<grid, 128, 0, st>>>(d_dibits, n, d_words, d_metric);
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

}  // namespace test

}  // namespace dh
