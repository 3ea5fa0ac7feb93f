// pocsag.cu — K3/K4 for POCSAG: 32-bit frame-sync search, BCH(31,21) + parity, message assembly; one warp per
// channel (sm_100a).
//
// Replaces Digiham::Pocsag::{SyncPhase,CodewordPhase}::process, Codeword::parse and Message
// (reference src/pocsag_decoder/pocsag_phase.cpp:10-92, codeword.cpp:9-55, message.cpp:7-73, bch_31_21.c).
// Input is one byte per bit (0/1) as FskDemodulator(40, invert) emits it (examples/pocsag-decoder.sh:19-21).
// Output is the decoder's byte stream: one line `address:N;message:TEXT\n` per completed alphanumeric message
// (the reference only ever creates messages for function 1 and 3 and only fills function 3, message.cpp:26-71).
//
// Warp-parallel pieces: the sync correlator tests 32 bit offsets per step (bit-plane by __ballot_sync, window by
// __funnelshift_r, XOR + __popc against the frame sync word 0x7CD215D8); a codeword is gathered with one ballot.
#include "decoder_ops.hpp"
#include "test_hooks.hpp"

#define DH_TABLES_NO_HOST_ARRAYS
#include "tables.inc"

#include <cstring>

namespace dh {

constexpr int kPocsagCarryCap = 48;
constexpr int kCodeword = 32;
constexpr int kMaxMessage = 80;   // MAX_MESSAGE_LENGTH (message.hpp:8)

struct PocsagState {
    int carry_len;
    int phase;            // 0 = SyncPhase, 1 = CodewordPhase
    int syncCount;
    int codewordCounter;
    int has_msg;
    int msg_type;
    int msg_pos;
    uint32_t msg_address;
    uint8_t content[kMaxMessage];
};

#ifdef __CUDACC__
namespace {

__constant__ uint32_t c_bch_lut[1024] = DH_BCH_31_21_LUT_INIT;
__constant__ uint32_t c_bch_h[10] = DH_BCH_31_21_H_INIT;

// frame sync word, transmitted MSB first (pocsag_phase.hpp:15); plane bit i = i-th received bit
constexpr uint32_t kFsc = 0x7CD215D8u;
// bch_31_21 (bch_31_21.c:521-561) on the 31 code bits of a codeword (parity bit already shifted out): syndrome ->
// direct LUT of the 1- and 2-bit error patterns; false when the syndrome is not in the table
__device__ __forceinline__ bool fec_bch31(uint32_t& payload) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 10; k++) s = (s << 1) | parity32(c_bch_h[k] & payload);
    if (s == 0) return true;
    const uint32_t e = c_bch_lut[s];
    payload ^= e;
    return e != 0;
}

__host__ __device__ constexpr uint32_t bit_reverse(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
constexpr uint32_t kFscPlane = bit_reverse(kFsc);
constexpr uint32_t kIdle = 0x7A89C197u;   // idle codeword (codeword.hpp:22)

struct PCtx {
    PocsagState st;   // scalar part; content lives in shared memory
    DecWriter w;
    uint8_t* content;
    int lane;
};

// Message::serialize (message.cpp:16-24) with the default StringSerializer: keys in std::map order
__device__ void serialize_message(PCtx& c) {
    if (!c.st.has_msg || c.st.msg_pos == 0) return;
    if (c.lane == 0) {
        uint8_t* o = c.w.out + c.w.out_len;
        uint32_t n = 0;
        const uint32_t room = c.w.out_cap - c.w.out_len;
        auto put = [&](uint8_t ch) {
            if (n < room) o[n] = ch;
            n++;
        };
        const char k1[] = "address:";
        for (int i = 0; i < 8; i++) put((uint8_t) k1[i]);
        char digits[10];
        int nd = 0;
        uint32_t a = c.st.msg_address;
        do {
            digits[nd++] = (char) ('0' + a % 10u);
            a /= 10u;
        } while (a);
        while (nd) put((uint8_t) digits[--nd]);
        const char k2[] = ";message:";
        for (int i = 0; i < 9; i++) put((uint8_t) k2[i]);
        for (int i = 0; i < kMaxMessage && c.content[i]; i++) put(c.content[i]);
        put('\n');
        // hand the record length to the other lanes through the scratch behind the message buffer
        c.content[kMaxMessage] = (uint8_t) (n & 0xFF);
        c.content[kMaxMessage + 1] = (uint8_t) (n >> 8);
    }
    __syncwarp();
    const uint32_t n = c.content[kMaxMessage] | ((uint32_t) c.content[kMaxMessage + 1] << 8);
    if (c.w.out_len + n <= c.w.out_cap) {
        c.w.out_len += n;
        // record boundary for callers with their own Serializer: {length of the rendered record, address digits}, so
        // that the host can slice {address, message} out of the byte stream without searching for separators
        uint32_t a = c.st.msg_address;
        uint8_t nd = 0;
        do {
            nd++;
            a /= 10u;
        } while (a);
        const uint8_t rec[2] = {(uint8_t) (n & 0xFF), (uint8_t) (n >> 8)};
        c.w.event(c.lane, 1, 0, nd, 0, rec, 2);
    } else {
        c.w.flags |= kFlagOutOverflow;
    }
    __syncwarp();
}

__device__ __forceinline__ void drop_message(PCtx& c) {
    c.st.has_msg = 0;
}

__device__ __forceinline__ void new_message(PCtx& c, uint32_t address, int type) {
    c.st.has_msg = 1;
    c.st.msg_type = type;
    c.st.msg_pos = 0;
    c.st.msg_address = address;
    for (int i = c.lane; i < kMaxMessage; i += 32) c.content[i] = 0;
    __syncwarp();
}

// Message::append (message.cpp:26-71): only function 3 (7-bit characters, LSB first) ever stores anything
__device__ __forceinline__ void append_message(PCtx& c, uint32_t data20) {
    if (c.st.msg_type != 3) return;
    if (c.st.msg_pos + 20 < kMaxMessage * 7) {
        if (c.lane == 0) {
            int pos = c.st.msg_pos;
            for (int i = 0; i < 20; i++, pos++) {
                const uint32_t bit = (data20 >> (19 - i)) & 1u;
                c.content[pos / 7] |= (uint8_t) (bit << (pos % 7));
            }
        }
        __syncwarp();
        c.st.msg_pos += 20;
    }
}

constexpr int kPWarps = 4;

__global__ void __launch_bounds__(kPWarps * 32) pocsag_kernel(const __grid_constant__ DecIo io, PocsagState* states) {
    __shared__ uint8_t s_content[kPWarps][kMaxMessage + 16];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kPWarps + warp;
    if (ch >= io.channels) return;

    PCtx c;
    c.lane = lane;
    c.content = s_content[warp];
    c.st = states[ch];
    for (int i = lane; i < kMaxMessage; i += 32) c.content[i] = states[ch].content[i];
    c.w.out = io.out + (size_t) ch * io.out_cap;
    c.w.ev = io.ev + (size_t) ch * io.ev_cap;
    c.w.out_len = io.out_len[ch];
    c.w.ev_len = io.ev_len[ch];
    c.w.out_cap = io.out_cap;
    c.w.ev_cap = io.ev_cap;
    c.w.flags = 0;
    __syncwarp();

    uint8_t* row = io.sym + (size_t) ch * io.sym_pitch;
    const int carry_len = c.st.carry_len;
    const uint8_t* stream = row + (io.carry_cap - carry_len);
    const int T = carry_len + (int) min((unsigned long long) io.nsym[ch], io.sym_pitch - io.carry_cap);
    int pos = 0;

    for (;;) {
        if (c.st.phase == 0) {
            // SyncPhase (pocsag_phase.cpp:18-28): more than 32 bits buffered, window at the read pointer
            const int avail = T - pos - kCodeword;
            if (avail <= 0) break;
            const int i0 = pos + lane;
            const uint8_t v0 = i0 < T ? stream[i0] : 0;
            const uint8_t v1 = i0 + 32 < T ? stream[i0 + 32] : 0;
            const uint32_t a = __ballot_sync(0xffffffffu, v0 & 1);
            const uint32_t b = __ballot_sync(0xffffffffu, v1 & 1);
            const uint32_t wnd = __funnelshift_r(a, b, lane);
            const bool hit = lane < avail && __popc(wnd ^ kFscPlane) <= 3;
            const uint32_t hits = __ballot_sync(0xffffffffu, hit);
            if (hits) {
                pos += (__ffs(hits) - 1) + kCodeword;   // the sync word is consumed
                c.st.phase = 1;
                c.st.syncCount = 1;
                c.st.codewordCounter = 0;
                c.st.has_msg = 0;
            } else {
                pos += min(32, avail);
            }
        } else {
            // CodewordPhase (pocsag_phase.cpp:38-92)
            if (T - pos <= kCodeword) break;
            const uint32_t plane = __ballot_sync(0xffffffffu, stream[pos + lane] != 0);   // input[i] && 1
            if (c.st.codewordCounter >= 16) {
                if (__popc(plane ^ kFscPlane) <= 3) {
                    if (c.st.syncCount++ > 2) c.st.syncCount = 2;
                } else {
                    if (c.st.syncCount-- < 0) {
                        serialize_message(c);
                        c.st.phase = 0;   // back to SyncPhase without consuming anything
                        c.st.has_msg = 0;
                        continue;
                    }
                }
                pos += kCodeword;
                c.st.codewordCounter = 0;
            } else {
                uint32_t cw = __brev(plane);   // first received bit = MSB
                uint32_t payload = cw >> 1;
                bool ok = fec_bch31(payload);
                cw = (cw & 1u) | (payload << 1);
                if (ok && parity32(cw)) ok = false;
                if (ok) {
                    if (cw == kIdle) {
                        serialize_message(c);
                        drop_message(c);
                    } else if ((cw >> 31) == 0) {
                        serialize_message(c);
                        drop_message(c);
                        const int type = (cw >> 11) & 3;
                        if (type == 1 || type == 3) {
                            const uint32_t address = (((cw >> 13) & 0x3FFFFu) << 3) | (uint32_t) (c.st.codewordCounter / 2);
                            new_message(c, address, type);
                        }
                    } else if (c.st.has_msg) {
                        append_message(c, (cw >> 11) & 0xFFFFFu);
                    }
                } else {
                    drop_message(c);
                }
                pos += kCodeword;
                c.st.codewordCounter++;
            }
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
    __syncwarp();
    for (int i = lane; i < kMaxMessage; i += 32) states[ch].content[i] = c.content[i];
}

}  // namespace
#endif  // __CUDACC__

namespace {

void pocsag_init_states(void* host_states, uint32_t count) {
    std::memset(host_states, 0, (size_t) count * sizeof(PocsagState));
}
// shortest message = address + one message codeword (64 bits) -> about 27 bytes of text
uint32_t pocsag_out_bytes(size_t max_syms) { return (uint32_t) ((max_syms + kPocsagCarryCap) / 2 + 256); }
// one record-boundary event per serialised message (address + at least one message codeword = 64 bits)
uint32_t pocsag_events(size_t max_syms) { return (uint32_t) ((max_syms + kPocsagCarryCap) / 64 + 4); }

int pocsag_launch(const DecIo& io, void* d_states, const uint8_t*, cudaStream_t stream) {
    const unsigned grid = (io.channels + kPWarps - 1) / kPWarps;
    pocsag_kernel<<<grid, kPWarps * 32, 0, stream>>>(io, static_cast<PocsagState*>(d_states));
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

const ProtoOps kPocsagOps = {"pocsag", sizeof(PocsagState), kPocsagCarryCap, pocsag_init_states, pocsag_out_bytes,
                             pocsag_events, pocsag_launch, make_pocsag_replay};

}  // namespace

const ProtoOps* pocsag_ops() { return &kPocsagOps; }

// ---- device-level test hook (dh_test_fec code 7) -------------------------------------------------------------------
namespace {

__global__ void pocsag_test_bch_kernel(uint32_t* words, uint8_t* ok, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w = words[i];
    const bool r = fec_bch31(w);
    words[i] = w;
    ok[i] = r ? 1 : 0;
}

}  // namespace

namespace test {

int pocsag_bch(uint32_t* d_words, uint8_t* d_ok, uint32_t n, cudaStream_t st) {
    if (n == 0) return DH_OK;
    pocsag_test_bch_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_words, d_ok, n);
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

}  // namespace test

}  // namespace dh
