// nxdn.cu — NXDN decoder bank: frame sync, LICH, SACCH superframes, FACCH1, voice extraction; one warp per channel
// (sm_100a).  SURVEY.md §8(f) rank 1.
//
// Device side replaces Digiham::Nxdn::{SyncPhase,FramedPhase}::process (reference
// src/nxdn_decoder/nxdn_phase.cpp:19-171), Scrambler (scrambler.cpp:8-25), Lich::parse (lich.cpp:5-33),
// Sacch::parse / SacchSuperframeCollector (sacch.cpp:24-134), Facch1::parse (facch1.cpp:8-75) and
// Trellis::decode (trellis.cpp:29-101, K5 in viterbi.cuh).  The metadata strings of Nxdn::MetaCollector
// (nxdn_meta.cpp:6-76) are replayed on the host from event records (meta_replay.cu).
//
// Contract quirks that are reproduced on purpose:
//   * punctured code bits are fed to the Viterbi decoder as hard zeros (sacch.cpp:56-67, facch1.cpp:45-56), so
//     clean FACCH1 blocks frequently fail their CRC — same failures here;
//   * a frame sync hit in SyncPhase does not consume the sync (nxdn_phase.cpp:23-24): FramedPhase re-checks it;
//   * a TX_RELEASE FACCH1 returns to SyncPhase WITHOUT consuming its own 72 symbols (nxdn_phase.cpp:149-152);
//   * the LICH of the last valid frame is kept when a new LICH fails its parity (nxdn_phase.cpp:64-69).
#include "decoder_ops.hpp"
#include "test_hooks.hpp"
#include "viterbi.cuh"
#include "crc_par.cuh"

#include <cstring>

namespace dh {

constexpr int kNxCarryCap = 208;
constexpr int kNxFrame = 192, kNxSync = 10;

enum : uint8_t {
    kNxEvSync = 1,    // setSync("voice")
    kNxEvSacch = 2,   // setFromSacch: a = call type, data = source (2, big endian), destination (2)
    kNxEvReset = 3,   // MetaCollector::reset()
};

struct NxdnState {
    int carry_len;
    int phase;            // 0 = SyncPhase, 1 = FramedPhase
    int syncCount;
    int lich;             // -1 = none yet, else the 7-bit LICH
    uint32_t sacch[4];    // 18 superframe bits of each collected fragment
    int sacch_mask;       // collected[i] != nullptr
    // mirror of Nxdn::MetaCollector, used to drop calls that cannot change it
    int m_sync, m_type, m_src, m_dst;
};

#ifdef __CUDACC__
namespace {

// frame sync {3,0,3,1,3,3,1,1,2,1} as dibit bit planes, symbol i -> bit i (nxdn_phase.cpp:16-17)
__host__ __device__ constexpr uint32_t fsw_plane(int which) {
    const int fsw[10] = {3, 0, 3, 1, 3, 3, 1, 1, 2, 1};
    uint32_t p = 0;
    for (int i = 0; i < 10; i++) p |= (uint32_t) ((which ? (fsw[i] >> 1) : fsw[i]) & 1) << i;
    return p;
}
constexpr uint32_t kFswHi = fsw_plane(1), kFswLo = fsw_plane(0);

__device__ __forceinline__ bool is_fsw(uint32_t hi, uint32_t lo) {
    const uint32_t m = 0x3FFu;
    return __popc((hi ^ kFswHi) & m) + __popc((lo ^ kFswLo) & m) <= 2;
}

// scrambler output for the 182 symbols behind the frame sync (scrambler.cpp:8-25): bit i of word i / 32
struct PnTable {
    uint32_t w[6];
};
__host__ __device__ constexpr PnTable make_pn() {
    PnTable t = {{0, 0, 0, 0, 0, 0}};
    unsigned sr = 0xE4;   // 0b011100100
    for (int i = 0; i < 182; i++) {
        const unsigned wb = sr & 1u;
        t.w[i >> 5] |= wb << (i & 31);
        const unsigned fb = ((sr >> 4) & 1u) ^ wb;
        sr = ((sr & 0x1FEu) >> 1) | (fb << 8);
    }
    return t;
}
__constant__ PnTable c_nx_pn = make_pn();

// Sacch::check_crc / Facch1::check_crc as warp-parallel table look-ups (crc_par.cuh)
__constant__ CrcTable<26> c_crc6 = make_crc_table<1, 26>();
__constant__ CrcTable<80> c_crc12 = make_crc_table<2, 80>();
struct NxdnCrcTables {
    uint16_t t6[26], t12[80];
};

struct NCtx {
    const NxdnCrcTables* crc;
    NxdnState st;
    DecWriter w;
    uint8_t* body;      // 182 descrambled dibits of the frame (shared memory)
    uint8_t* scratch;   // 96 bytes of per-warp scratch (shared memory)
    int lane;
};

__device__ __forceinline__ void meta_reset(NCtx& c) {
    NxdnState& s = c.st;
    if (s.m_sync || s.m_type || s.m_src || s.m_dst) {
        c.w.event(c.lane, kNxEvReset, 0);
        s.m_sync = s.m_type = s.m_src = s.m_dst = 0;
    }
}

// bit `pos` of a dibit-per-byte array, high bit first (sacch.cpp:50, facch1.cpp:39)
__device__ __forceinline__ uint32_t dibit_bit(const uint8_t* d, int pos) { return (d[pos >> 1] >> (1 - (pos & 1))) & 1u; }

// Sacch::parse (sacch.cpp:24-43): de-interleave 12 x 5, re-insert the punctured positions ((i + 1) % 6 == 0) as
// zeros, Viterbi over 36 steps, 6-bit CRC over 26 bits.  Returns true and the first 32 decoded bits.
__device__ bool parse_sacch(NCtx& c, const uint8_t* in, uint32_t& word) {
    uint8_t* dib = c.scratch;
    for (int p = c.lane; p < 36; p += 32) {
        uint32_t v = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int j = 2 * p + h;
            uint32_t bit = 0;
            if ((j + 1) % 6 != 0) {
                const int q = j - j / 6;                 // index in the punctured stream
                bit = dibit_bit(in, (q % 12) * 5 + q / 12);
            }
            v = (v << 1) | bit;
        }
        dib[p] = (uint8_t) v;
    }
    __syncwarp();
    uint32_t words[2];
    viterbi<36, true>(dib, c.lane, words);
    __syncwarp();
    const uint32_t crc = crc_parallel<26>(words, c.crc->t6, c_crc6.c, c.lane);
    if ((words[0] & 0x3Fu) != crc) return false;
    word = words[0];
    return true;
}

// Facch1::parse (facch1.cpp:8-26): de-interleave 16 x 9, punctured positions ((i - 1) % 4 == 0), Viterbi over 96
// steps, 12-bit CRC over 80 bits.  Returns the message type or -1.
__device__ int parse_facch1(NCtx& c, const uint8_t* in) {
    uint8_t* dib = c.scratch;
    for (int p = c.lane; p < 96; p += 32) {
        uint32_t v = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int j = 2 * p + h;
            uint32_t bit = 0;
            if ((j & 3) != 1) {
                const int q = j - (j + 2) / 4;
                bit = dibit_bit(in, (q % 16) * 9 + q / 16);
            }
            v = (v << 1) | bit;
        }
        dib[p] = (uint8_t) v;
    }
    __syncwarp();
    uint32_t words[3];
    viterbi<96, true>(dib, c.lane, words);
    __syncwarp();
    const uint32_t crc = crc_parallel<80>(words, c.crc->t12, c_crc12.c, c.lane);
    const uint32_t to_check = (words[2] >> 4) & 0xFFFu;   // bits 80..91
    if (to_check != crc) return -1;
    return (int) ((words[0] >> 24) & 0x3Fu);
}

// SacchSuperframeCollector::push / isComplete / getSuperframe (sacch.cpp:96-134) + MetaCollector::setFromSacch
__device__ void collect_sacch(NCtx& c, uint32_t word) {
    NxdnState& s = c.st;
    const int index = (int) ((word >> 30) ^ 3u);
    if (index > 0 && !((s.sacch_mask >> (index - 1)) & 1)) return;   // fragment before it is missing
    const uint32_t fragment = (word >> 6) & 0x3FFFFu;                  // bits 8..25
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k == index) s.sacch[k] = fragment;                         // static indices: the state stays in registers
    }
    s.sacch_mask |= 1 << index;
    if (s.sacch_mask != 0xF) return;
    // 72 bits = 4 x 18, MSB first; only bytes 0, 2, 3..6 are read (sacch.cpp:140-154)
    const unsigned long long top = ((unsigned long long) s.sacch[0] << 46) | ((unsigned long long) s.sacch[1] << 28) |
                                   ((unsigned long long) s.sacch[2] << 10) | (s.sacch[3] >> 8);
    const int message_type = (int) ((top >> 56) & 0x3F);
    if (message_type == 0x01) {   // NXDN_MESSAGE_TYPE_VCALL
        const int call_type = (int) ((top >> 45) & 7);
        const int type = call_type == 1 ? 1 : (call_type == 4 ? 2 : 0);   // conference / individual / ""
        const int src = (int) ((top >> 24) & 0xFFFF);
        const int dst = (int) ((top >> 8) & 0xFFFF);
        if (type != s.m_type || src != s.m_src || dst != s.m_dst) {
            const uint8_t d[4] = {(uint8_t) (src >> 8), (uint8_t) src, (uint8_t) (dst >> 8), (uint8_t) dst};
            c.w.event(c.lane, kNxEvSacch, 0, (uint8_t) type, 0, d, 4);
            s.m_type = type;
            s.m_src = src;
            s.m_dst = dst;
        }
    }
    s.sacch_mask = 0;
}

// FramedPhase::process (nxdn_phase.cpp:45-171) on the 192 staged symbols fr[].  Returns the number of symbols
// consumed; `to_sync` tells that the phase fell back to SyncPhase.
__device__ int nxdn_frame(NCtx& c, const uint8_t* fr, bool& to_sync) {
    NxdnState& s = c.st;
    const int lane = c.lane;
    to_sync = false;
    {
        const uint8_t v = lane < kNxSync ? fr[lane] : 0;
        const uint32_t hi = __ballot_sync(0xffffffffu, (v >> 1) & 1);
        const uint32_t lo = __ballot_sync(0xffffffffu, v & 1);
        if (is_fsw(hi, lo)) {
            if (++s.syncCount > 6) s.syncCount = 6;
        } else if (--s.syncCount < 0) {
            meta_reset(c);
            to_sync = true;
            return 0;
        }
    }
    // descramble everything behind the sync: the high bit of the dibit is inverted where the PN output is 1
    for (int i = lane; i < kNxFrame - kNxSync; i += 32) {
        const uint32_t wb = (c_nx_pn.w[i >> 5] >> (i & 31)) & 1u;
        c.body[i] = (uint8_t) ((fr[kNxSync + i] & 3u) ^ (wb << 1));
    }
    __syncwarp();

    // Lich::parse (lich.cpp:5-33)
    {
        const uint32_t bits = __ballot_sync(0xffffffffu, lane < 8 && ((c.body[lane] >> 1) & 1));   // bit i = lich_bits[i]
        const uint32_t check = (bits ^ (bits >> 1) ^ (bits >> 2) ^ (bits >> 3)) & 1u;
        if (((bits >> 7) & 1u) == check) s.lich = (int) (__brev(bits & 0x7Fu) >> 25);             // bit i -> 6 - i
    }
    if (s.lich < 0) return kNxFrame;
    const int rf = (s.lich >> 5) & 3, functional = (s.lich >> 3) & 3, option = (s.lich >> 1) & 3;
    if (rf == 0 /* RCCH */ || functional == 1 /* UDCH */) return kNxFrame;

    if (functional == 2 /* SACCH superframe */) {
        uint32_t word;
        if (parse_sacch(c, c.body + 8, word)) collect_sacch(c, word);
    }
    for (int i = 0; i < 2; i++) {
        const uint8_t* half = c.body + 38 + 72 * i;
        if ((option >> (1 - i)) & 1) {
            if (s.syncCount >= 1) {
                if (!s.m_sync) {
                    c.w.event(lane, kNxEvSync, 0);
                    s.m_sync = 1;
                }
                if (c.w.out_len + 18 <= c.w.out_cap) {
                    if (lane < 18) {
                        const uint8_t* p = half + 4 * lane;
                        c.w.out[c.w.out_len + lane] =
                            (uint8_t) (((p[0] & 3u) << 6) | ((p[1] & 3u) << 4) | ((p[2] & 3u) << 2) | (p[3] & 3u));
                    }
                    c.w.out_len += 18;
                } else {
                    c.w.flags |= kFlagOutOverflow;
                }
            }
        } else {
            const int type = parse_facch1(c, half);
            if (type == 0x08) {   // NXDN_MESSAGE_TYPE_TX_RELEASE: back to SyncPhase, this block is not consumed
                meta_reset(c);
                to_sync = true;
                return kNxSync + 8 + 30 + 72 * i;
            }
        }
    }
    return kNxFrame;
}

constexpr int kNWarps = 4;

__global__ void __launch_bounds__(kNWarps * 32) nxdn_kernel(const __grid_constant__ DecIo io, NxdnState* states) {
    __shared__ __align__(16) uint8_t s_fr[kNWarps][kNxFrame + 8];
    __shared__ __align__(16) uint8_t s_body[kNWarps][kNxFrame];
    __shared__ __align__(16) uint8_t s_scratch[kNWarps][96];
    __shared__ NxdnCrcTables s_crc;
    for (int i = threadIdx.x; i < 80; i += kNWarps * 32) {
        if (i < 26) s_crc.t6[i] = c_crc6.t[i];
        s_crc.t12[i] = c_crc12.t[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kNWarps + warp;
    if (ch >= io.channels) return;

    NCtx c;
    c.st = states[ch];
    c.lane = lane;
    c.crc = &s_crc;
    c.body = s_body[warp];
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
            // SyncPhase (nxdn_phase.cpp:19-32): more than 10 symbols buffered, sync at the read pointer
            const int avail = T - pos - kNxSync;
            if (avail <= 0) break;
            const int i0 = pos + lane;
            const uint8_t v0 = i0 < T ? stream[i0] : 0;
            const uint8_t v1 = i0 + 32 < T ? stream[i0 + 32] : 0;
            const uint32_t a_hi = __ballot_sync(0xffffffffu, (v0 >> 1) & 1);
            const uint32_t a_lo = __ballot_sync(0xffffffffu, v0 & 1);
            const uint32_t b_hi = __ballot_sync(0xffffffffu, (v1 >> 1) & 1);
            const uint32_t b_lo = __ballot_sync(0xffffffffu, v1 & 1);
            const bool hit = lane < avail && is_fsw(__funnelshift_r(a_hi, b_hi, lane), __funnelshift_r(a_lo, b_lo, lane));
            const uint32_t hits = __ballot_sync(0xffffffffu, hit);
            if (hits) {
                pos += __ffs(hits) - 1;
                // fresh FramedPhase (nxdn_phase.hpp:31-43): the sync itself is not consumed
                c.st.phase = 1;
                c.st.syncCount = 0;
                c.st.lich = -1;
                c.st.sacch_mask = 0;
            } else {
                pos += min(32, avail);
            }
        } else {
            if (T - pos <= kNxFrame) break;
            const uint8_t* fr = stage_symbols(s_fr[warp], stream + pos, kNxFrame, lane);
            __syncwarp();
            bool to_sync;
            pos += nxdn_frame(c, fr, to_sync);
            if (to_sync) c.st.phase = 0;
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

void nxdn_init_states(void* host_states, uint32_t count) {
    std::memset(host_states, 0, (size_t) count * sizeof(NxdnState));
    NxdnState* s = static_cast<NxdnState*>(host_states);
    for (uint32_t i = 0; i < count; i++) s[i].lich = -1;
}
// at most 36 voice bytes per 192-symbol frame; a TX_RELEASE can end a frame after 48 symbols without output
uint32_t nxdn_out_bytes(size_t max_syms) { return (uint32_t) (36 * ((max_syms + kNxCarryCap) / kNxFrame + 2)); }
// per frame at most: sync, sacch, reset
uint32_t nxdn_events(size_t max_syms) { return (uint32_t) (3 * ((max_syms + kNxCarryCap) / 48 + 2) + 8); }

int nxdn_launch(const DecIo& io, void* d_states, const uint8_t*, cudaStream_t stream) {
    const unsigned grid = (io.channels + kNWarps - 1) / kNWarps;
    nxdn_kernel<<<grid, kNWarps * 32, 0, stream>>>(io, static_cast<NxdnState*>(d_states));
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

const ProtoOps kNxdnOps = {"nxdn", sizeof(NxdnState), kNxCarryCap, nxdn_init_states, nxdn_out_bytes, nxdn_events,
                           nxdn_launch, make_nxdn_replay};

}  // namespace

const ProtoOps* nxdn_ops() { return &kNxdnOps; }

// ---- device-level test hook (dh_test_viterbi variants 2 / 3) ------------------------------------------------------
namespace {

template <int STEPS>
__global__ void nxdn_test_viterbi_kernel(const uint8_t* dibits, uint32_t n, uint32_t* words, uint32_t* metric) {
    constexpr int NW = (STEPS + 31) / 32;
    __shared__ uint8_t stage[4][96];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * 4 + wib;
    if (i >= n) return;
    for (int k = lane; k < STEPS; k += 32) stage[wib][k] = dibits[(size_t) i * STEPS + k];
    __syncwarp();
    uint32_t w[NW];
    const uint32_t m = viterbi<STEPS, true>(stage[wib], lane, w);
    if (lane == 0) {
        for (int k = 0; k < NW; k++) words[(size_t) i * NW + k] = w[k];
        metric[i] = m;
    }
}

}  // namespace

namespace test {

int nxdn_viterbi(int steps, const uint8_t* d_dibits, uint32_t n, uint32_t* d_words, uint32_t* d_metric, cudaStream_t st) {
    DH_REQUIRE(steps == 36 || steps == 96, DH_E_INVALID, "dh_test_viterbi: NXDN decodes 36 or 96 steps");
    if (n == 0) return DH_OK;
    const unsigned grid = (n + 3) / 4;
    if (steps == 36) nxdn_test_viterbi_kernel<36><<<grid, 128, 0, st>>>(d_dibits, n, d_words, d_metric);
    else nxdn_test_viterbi_kernel<96><<<grid, 128, 0, st>>>(d_dibits, n, d_words, d_metric);
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

}  // namespace test

}  // namespace dh
