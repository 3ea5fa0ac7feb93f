// demod.cu — K2: 2-/4-level FSK slicer with variance-minimum symbol timing (dh_demod_*), sm_100a.
//
// Replaces Digiham::Fsk::GfskDemodulator / FskDemodulator (reference src/gfsk_demodulator/gfsk_demodulator.cpp:18-122,
// src/fsk_demodulator/fsk_demodulator.cpp:19-112) for N channels at once, bit-exactly.
//
// What the reference does per symbol, restated (there is NO Gardner loop and no interpolator):
//   * read sps samples at the read pointer: `sum` over the middle third, `volume_sum` over all, in order, fp32;
//   * advance by sps + variance_offset (offset decided at the end of the previous 100-symbol block, so it shifts
//     every window of the current block except the first one);
//   * every 100 symbols: per sample phase i, fp32 ordered total -> mean (float divide, widened) -> fp64 ordered
//     sum of squared deviations -> /100; first strict minimum decides a +-1 sample nudge for the next block;
//   * push volume_sum/sps into a 100-entry ring, min/max over the ring (max starts at FLT_MIN), thresholds in
//     mixed float/double arithmetic, slice.
//
// GPU shape: the only sequential dependency is the +-1 nudge from one 100-symbol block to the next, so a group of
// G lanes owns one channel and walks its blocks in order: G = 10 at sps 10 (three channels per warp, compile-time
// fast path), G = 20 at sps 20 / 40 (fast paths with 128-bit window loads), G = 16 / 32 for any other sps (generic
// kernel).  Inside a block the window positions are known up front: lanes own symbols for the window sums, then
// the ring min/max becomes (prefix over this block) x (suffix over the previous block) computed with shuffle
// scans, then lanes own sample phases for the variance search.  Samples are staged block-wise into shared memory
// with 16-byte cp.async copies — from the bank's own work row (a producer kernel wrote the chunk there) or straight
// from the caller's rows (DemodParams::ext).  Partially filled blocks are carried: the unconsumed tail of every
// channel is moved right-aligned in front of the position where the next chunk starts, in the other of two
// alternating work-row sets.
//
// Small banks (up to 1024 channels by default, dh_demod_set_split) run the same arithmetic as three kernels instead:
// only the variance search walks the blocks in order, window sums and slicing run one lane per symbol / one lane
// group per block (see "Split form of K2" below); with few channels the block chain is the whole step.
#include "common.cuh"

#include <cstring>

#include <cfloat>
#include <type_traits>
#include <cstdlib>
#include <new>

namespace {

constexpr int kBlockSyms = 100;  // VARIANCE_SYMBOLS == VOLUME_RB_SIZE == 100 (include/gfsk_demodulator.hpp:5-6)
// 2 warps: 6 channels per CTA at sps = 10 (39 KB smem, 5 CTAs per SM).  4-warp CTAs (78 KB, 2 per SM) were measured
// slower on B200 (0.42 vs 0.34 ms at 4096 channels: 342 CTAs are 1.15 waves of 2 x 148).
constexpr int kThreads = 64;
constexpr int kCarrySlack = 16;  // carry_cap = 100 * sps + kCarrySlack
// dh_demod_set_split default of new banks (environment DH_DEMOD_SPLIT overrides it): -1 = by bank and call size.
// Measured on B200 (profiles/r02_split_small_banks.txt, DMR pipe, 48000 samples per call): the split schedule is
// 42 % faster per step at 1 channel, 39 % at 32, 30 % at 256, 22 % at 512, 10 % at 1024, equal at 2048 and 11 %
// slower at 4096, where the one-kernel form has enough channels in flight and the split pays for reading the
// samples twice.
constexpr int kSplitDefault = -1;
constexpr uint32_t kSplitAutoChannels = 1024;  // auto: banks up to this size ...
constexpr size_t kSplitAutoBlocks = 8;         // ... and calls that span at least this many 100-symbol blocks

struct ChannelState {
    float vol_prev[kBlockSyms];  // volume ring content written by the previous block (zeros at power-on)
    int vo;                      // variance_offset pending for the current block
    int j_done;                  // symbols of the current block already emitted
    int carry_len;               // samples kept in front of the work row
    int pad_;
};

struct DemodParams {
    const float* work;           // [channels][pitch] rows read by this call; chunk sample 0 at column carry_cap
    float* work_next;            // rows the next call reads: receives the carried tail
    const float* ext;            // non-null: the chunk lives in the CALLER's rows (16-byte aligned), only the carried
    unsigned long long ext_pitch;  //         tail comes from `work`; null: the chunk follows the tail inside `work`
    unsigned long long pitch;
    uint8_t* sym;                // [channels][sym_pitch]
    unsigned long long sym_pitch;
    uint32_t* nsym;              // [channels] symbols emitted by this call
    ChannelState* state;
    int channels;
    int n;                       // new samples per channel
    int sps;
    int lo, hi;                  // evaluation window [lo, hi)
    int four_level;
    int invert;
    int carry_cap;
    int samples_cap;             // floats reserved per group for staged samples
    int group_floats;            // floats of shared memory per group
    // split mode only (demod_search_kernel -> demod_volume_kernel -> demod_slice_kernel)
    ChannelState* state_out;     // state the NEXT call reads (the two state sets alternate like the work rows)
    int2* rec;                   // [channels][rec_pitch] {first window, pending nudge} of every block of this call
    int4* hdr;                   // [channels] {full blocks, symbols of the trailing partial block, j_done at entry,
                                 //             carry_len at entry}
    float2* va;                  // [channels][rec_pitch * 100] {volume, average} of every symbol of this call
    int rec_pitch;
    int search_dbuf;             // search kernel: two sample buffers per group (next block staged during the search)
};

__device__ __forceinline__ float min_lt(float cur, float v) { return v < cur ? v : cur; }
__device__ __forceinline__ float max_gt(float cur, float v) { return v > cur ? v : cur; }

// x / c for the divisors of the fast paths, as fl32(fl64(x) * fl64(1/c)): bit-identical to the float division for
// EVERY float x when c is 6, 10, 14, 20, 40 or 100 (exhaustive proof: tests/test_host_logic.py::
// test_division_by_constant_exhaustive), and three instructions instead of the IEEE division sequence.
__device__ __forceinline__ float div_by_const(float x, double reciprocal) {
    return __double2float_rn(__dmul_rn((double) x, reciprocal));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dh::smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dh::smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// G lanes own one channel; a warp carries 32 / G channels (3 at G = 10, lanes 30 and 31 idle).  Lane gl owns the
// kChunk consecutive symbols [gl * kChunk, (gl + 1) * kChunk) of every 100-symbol block.
template <int G>
struct Group {
    static constexpr int kPerWarp = 32 / G;
    static constexpr int kChunk = (kBlockSyms + G - 1) / G;
};

// value held by the lane d positions below / above inside the group (own value when there is none)
__device__ __forceinline__ float group_up(unsigned gmask, float v, int d, int gl) {
    const int lane = threadIdx.x & 31;
    return __shfl_sync(gmask, v, gl >= d ? lane - d : lane);
}
template <int G>
__device__ __forceinline__ float group_down(unsigned gmask, float v, int d, int gl) {
    const int lane = threadIdx.x & 31;
    return __shfl_sync(gmask, v, gl + d < G ? lane + d : lane);
}

// exclusive scans of per-lane (min, max) totals across the group: up = over lower lanes, down = over higher lanes
template <int G>
__device__ __forceinline__ void exclusive_up(unsigned gmask, int gl, float& mn, float& mx) {
    float tmn = mn, tmx = mx;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const float a = group_up(gmask, tmn, d, gl);
        const float b = group_up(gmask, tmx, d, gl);
        if (gl >= d) {
            tmn = min_lt(tmn, a);
            tmx = max_gt(tmx, b);
        }
    }
    mn = group_up(gmask, tmn, 1, gl);
    mx = group_up(gmask, tmx, 1, gl);
    if (gl == 0) {
        mn = FLT_MAX;
        mx = FLT_MIN;
    }
}
template <int G>
__device__ __forceinline__ void exclusive_down(unsigned gmask, int gl, float& mn, float& mx) {
    float tmn = mn, tmx = mx;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const float a = group_down<G>(gmask, tmn, d, gl);
        const float b = group_down<G>(gmask, tmx, d, gl);
        if (gl + d < G) {
            tmn = min_lt(tmn, a);
            tmx = max_gt(tmx, b);
        }
    }
    mn = group_down<G>(gmask, tmn, 1, gl);
    mx = group_down<G>(gmask, tmx, 1, gl);
    if (gl == G - 1) {
        mn = FLT_MAX;
        mx = FLT_MIN;
    }
}

// smn[q] / smx[q] = min / max(FLT_MIN, .) over the previous block's volumes of all symbols AFTER symbol
// gl * kChunk + q (empty range -> FLT_MAX / FLT_MIN): the part of the 100-entry ring the current block has not
// overwritten yet when its symbol j is sliced
template <int G>
__device__ __forceinline__ void suffix_of_previous(const float* pv, float* smn, float* smx, int gl, unsigned gmask) {
    constexpr int CH = Group<G>::kChunk;
    float mn = FLT_MAX, mx = FLT_MIN;
#pragma unroll
    for (int q = CH - 1; q >= 0; q--) {
        smn[q] = mn;
        smx[q] = mx;
        if (gl * CH + q < kBlockSyms) {
            mn = min_lt(mn, pv[q]);
            mx = max_gt(mx, pv[q]);
        }
    }
    exclusive_down<G>(gmask, gl, mn, mx);
#pragma unroll
    for (int q = 0; q < CH; q++) {
        smn[q] = min_lt(smn[q], mn);
        smx[q] = max_gt(smx[q], mx);
    }
}

// number of symbols of the block starting at P that can be processed with T samples visible: symbol j starts at
// P + j*sps + (j >= 1 ? vo : 0) and needs more than sps + 1 samples from there (gfsk_demodulator.cpp:21)
__device__ __forceinline__ int processable(int T, int P, int vo, int sps) {
    if (T - P < sps + 2) return 0;
    const int room = T - P - vo - sps - 2;
    return 1 + (room >= sps ? min(kBlockSyms - 1, room / sps) : 0);
}

// SPS > 0: compile-time samples per symbol (fast paths for 10, 20 and 40 = DMR/YSF/D-Star, NXDN, POCSAG at 48 kHz:
// unrolled windows, evaluation window lo/hi as constants, divisions by sps, hi - lo and 100 as reciprocal
// multiplications proven exact for every float), SPS == 0: run-time p.sps
__host__ __device__ constexpr int eval_lo(int sps) { return (2 * sps + 3) / 6; }   // == roundf(sps / 3.0f) for 10, 20, 40 (checked by the host)
__host__ __device__ constexpr int eval_hi(int sps) { return (4 * sps + 3) / 6; }   // == roundf(sps * 2 / 3.0f)

// MINB: minimum resident CTAs per SM the register allocation must allow (caps the registers per thread)
template <int G, int SPS, int THREADS, int MINB = 0>
__global__ void __launch_bounds__(THREADS, MINB) demod_kernel(const __grid_constant__ DemodParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int kPerWarp = Group<G>::kPerWarp;
    constexpr int CH = Group<G>::kChunk;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp_in_warp = lane / G;
    if (grp_in_warp >= kPerWarp) return;   // lanes that do not fill a whole group (30, 31 at G = 10)
    const int gl = lane - grp_in_warp * G;
    const int grp = warp * kPerWarp + grp_in_warp;
    const int ch = blockIdx.x * ((THREADS / 32) * kPerWarp) + grp;
    if (ch >= p.channels) return;          // a whole group leaves together; shuffles below only name the own group
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (grp_in_warp * G);

    const int sps = SPS > 0 ? SPS : p.sps;
    const int lo = SPS > 0 ? eval_lo(SPS) : p.lo;
    const int hi = SPS > 0 ? eval_hi(SPS) : p.hi;
    float* S = smem + (size_t) grp * p.group_floats;                 // staged samples of the current block
    double* var = reinterpret_cast<double*>(S + p.samples_cap);     // [sps] phase variances

    ChannelState* st = p.state + ch;
    int vo = st->vo;
    int j_done = st->j_done;
    const int carry_len = st->carry_len;
    const int j_first = gl * CH;           // first symbol of every block this lane owns
    float pv[CH], smn[CH], smx[CH];        // previous block's volumes and their exclusive suffix min / max
#pragma unroll
    for (int q = 0; q < CH; q++) pv[q] = j_first + q < kBlockSyms ? st->vol_prev[j_first + q] : 0.0f;
    suffix_of_previous<G>(pv, smn, smx, gl, gmask);

    const float* row = p.work + (size_t) ch * p.pitch;
    const int col0 = p.carry_cap - carry_len;   // column of logical stream index 0
    const int T = carry_len + p.n;              // samples visible to this call
    uint8_t* sym_row = p.sym + (size_t) ch * p.sym_pitch;
    const float fsps = (float) sps;
    const float fwin = (float) (hi - lo);

    // Logical stream = carried tail (in `work`) followed by the chunk: either right behind it in the same row
    // (a producer kernel wrote it there) or in the caller's own rows (`ext`), which are then read in place.
    const float* ext_row = p.ext ? p.ext + (size_t) ch * p.ext_pitch : nullptr;
    auto at = [&](int x) -> const float* {
        return (ext_row != nullptr && x >= carry_len) ? ext_row + (x - carry_len) : row + col0 + x;
    };
    // offset of sample P inside its 16-byte unit (0 for a block that starts in the tail of an `ext` call: that
    // block is copied element by element because it straddles two buffers)
    auto align_of = [&](int P) -> int {
        if (ext_row == nullptr) return (col0 + P) & 3;
        return P >= carry_len ? (P - carry_len) & 3 : 0;
    };
    // asynchronous staging of [P, P + m*sps + 2) with aligned 16-byte copies; sample P + x lands at S[a0 + x]
    auto stage = [&](int P, int m) {
        const int a0 = align_of(P);
        const int len = m * sps + 2;
        if (ext_row == nullptr) {
            const float4* src = reinterpret_cast<const float4*>(row + col0 + P - a0);
            float4* dst = reinterpret_cast<float4*>(S);
            const int nvec = (a0 + len + 3) >> 2;
            for (int v = gl; v < nvec; v += G) cp_async16(dst + v, src + v);
        } else if (P >= carry_len) {
            // whole vectors as long as they end inside the caller's row, single samples behind them (never read
            // past sample T - 1: the buffer is not ours)
            const float* base = ext_row + (P - carry_len) - a0;
            const int avail = T - P + a0;                       // floats from `base` to the end of the chunk
            const int want = a0 + len;
            const int nvec = min(want, avail) >> 2;
            const float4* src = reinterpret_cast<const float4*>(base);
            float4* dst = reinterpret_cast<float4*>(S);
            for (int v = gl; v < nvec; v += G) cp_async16(dst + v, src + v);
            for (int x = 4 * nvec + gl; x < min(want, avail); x += G) cp_async4(S + x, base + x);
        } else {
            for (int x = gl; x < len && P + x < T; x += G) cp_async4(S + x, at(P + x));
        }
        cp_async_commit();
    };

    int P = 0;         // logical index of the current block's first window
    int emitted = 0;
    int m = processable(T, P, vo, sps);
    if (m > j_done) {
        stage(P, m);
        cp_async_wait_all();
        __syncwarp(gmask);
    }
    while (m > j_done) {
        const int a0 = align_of(P);

        // window sums of the symbols this lane owns (gfsk_demodulator.cpp:28-35, 82-83, 88)
        float volr[CH], avgr[CH];
#pragma unroll
        for (int q = 0; q < CH; q++) {
            const int j = j_first + q;
            volr[q] = 0.0f;
            avgr[q] = 0.0f;
            if (j < m) {
                const float* w = S + a0 + j * sps + (j ? vo : 0);
                float sum = 0.0f, vsum = 0.0f;
                if (SPS > 0) {
                    constexpr int kLo = eval_lo(SPS > 0 ? SPS : 10), kHi = eval_hi(SPS > 0 ? SPS : 10);
                    if (SPS == 40 || (SPS == 20 && G == 20)) {
                        // Wide symbols: the lanes' windows are 200 floats apart (8 banks), so scalar loads collide five
                        // ways.  The windows share the same offset r inside a 16-byte unit ((a0 + vo) & 3; only symbol 0
                        // of a block, which is not shifted by the pending nudge, may differ), so they are fetched as 11
                        // aligned 128-bit loads and consumed in order from registers: a quarter of the load
                        // instructions and 2.5 x fewer wavefronts (3.34 -> 2.4 ms per 32768-channel launch).
                        const int r = (a0 + (j ? vo : 0)) & 3;
                        const float4* wa = reinterpret_cast<const float4*>(w - r);
                        constexpr int NV = (SPS + 6) / 4;   // vectors that cover r + SPS floats for every r <= 3
                        float x[4 * NV];
#pragma unroll
                        for (int v4 = 0; v4 < NV; v4++) {
                            const float4 t = wa[v4];
                            x[4 * v4] = t.x;
                            x[4 * v4 + 1] = t.y;
                            x[4 * v4 + 2] = t.z;
                            x[4 * v4 + 3] = t.w;
                        }
                        auto consume = [&](auto R) {
                            constexpr int kR = decltype(R)::value;
#pragma unroll
                            for (int i = 0; i < SPS; i++) {
                                const float v = x[i + kR];
                                if (i >= kLo && i < kHi) sum = __fadd_rn(sum, v);
                                vsum = __fadd_rn(vsum, v);
                            }
                        };
                        if (r == 0) consume(std::integral_constant<int, 0>());
                        else if (r == 1) consume(std::integral_constant<int, 1>());
                        else if (r == 2) consume(std::integral_constant<int, 2>());
                        else consume(std::integral_constant<int, 3>());
                    } else {
#pragma unroll
                    for (int i = 0; i < SPS; i++) {
                        const float v = w[i];
                        if (i >= kLo && i < kHi) sum = __fadd_rn(sum, v);
                        vsum = __fadd_rn(vsum, v);
                    }
                    }
                    volr[q] = div_by_const(vsum, 1.0 / (double) (SPS > 0 ? SPS : 1));
                    avgr[q] = kHi - kLo == 4 ? __fmul_rn(sum, 0.25f)   // / 4.0f, exact scaling
                                             : div_by_const(sum, 1.0 / (double) (kHi - kLo));
                } else {
                    for (int i = 0; i < sps; i++) {
                        const float v = w[i];
                        if (i >= lo && i < hi) sum = __fadd_rn(sum, v);
                        vsum = __fadd_rn(vsum, v);
                    }
                    volr[q] = __fdiv_rn(vsum, fsps);
                    avgr[q] = __fdiv_rn(sum, fwin);
                }
            }
        }

        // variance-minimum phase search over the 100 windows of a complete block (gfsk_demodulator.cpp:41-80)
        int vo_next = 0, P_next = 0, m_next = 0;
        const bool full = m == kBlockSyms;
        if (full) {
            for (int i = gl; i < sps; i += G) {
                const float* w0 = S + a0 + i;
                const float* wv = w0 + vo;          // windows 1..99 are shifted by the pending nudge
                float total = __fadd_rn(0.0f, w0[0]);
                if (SPS > 0) {
#pragma unroll 11
                    for (int k = 1; k < kBlockSyms; k++) total = __fadd_rn(total, wv[k * SPS]);
                } else {
                    for (int k = 1; k < kBlockSyms; k++) total = __fadd_rn(total, wv[k * sps]);
                }
                const double mean = (double) (SPS > 0 ? div_by_const(total, 1.0 / 100.0) : __fdiv_rn(total, 100.0f));
                double d = __dsub_rn(mean, (double) w0[0]);
                double dsum = __dadd_rn(0.0, __dmul_rn(d, d));
                if (SPS > 0) {
#pragma unroll 11
                    for (int k = 1; k < kBlockSyms; k++) {
                        d = __dsub_rn(mean, (double) wv[k * SPS]);
                        dsum = __dadd_rn(dsum, __dmul_rn(d, d));
                    }
                } else {
                    for (int k = 1; k < kBlockSyms; k++) {
                        d = __dsub_rn(mean, (double) wv[k * sps]);
                        dsum = __dadd_rn(dsum, __dmul_rn(d, d));
                    }
                }
                var[i] = __ddiv_rn(dsum, 100.0);
            }
            __syncwarp(gmask);
            double vmin = var[0];
            int vpos = 0;
            for (int i = 1; i < sps; i++) {
                const double v = var[i];
                if (v < vmin) {
                    vmin = v;
                    vpos = i;
                }
            }
            if (vmin <= 0 || vmin > 5000000) {
                // no decision
            } else if (vpos > 0 && vpos < sps / 2) {
                vo_next = +1;
            } else if (vpos >= sps / 2 && vpos < sps - 1) {
                vo_next = -1;
            }
            // the sample buffer is free from here on: start fetching the next block while this one is sliced
            P_next = P + kBlockSyms * sps + vo;
            m_next = processable(T, P_next, vo_next, sps);
            __syncwarp(gmask);
            if (m_next > 0) stage(P_next, m_next);
        }

        // Ring min/max right after symbol j was pushed = (this block's volumes 0..j) x (previous block's volumes
        // j+1..99).  Prefix part: running min/max over the own symbols on top of the exclusive scan over the lower
        // lanes; only valid symbols (j < m) take part.
        float emn = FLT_MAX, emx = FLT_MIN;
#pragma unroll
        for (int q = 0; q < CH; q++) {
            if (j_first + q < m) {
                emn = min_lt(emn, volr[q]);
                emx = max_gt(emx, volr[q]);
            }
        }
        exclusive_up<G>(gmask, gl, emn, emx);

        // calibrateAudio + slicing (gfsk_demodulator.cpp:88-104, 109-122)
        float rmn = emn, rmx = emx;
        uint8_t* const sym_out = sym_row + (emitted - j_done + j_first);   // this lane's symbol q goes to sym_out[q]
#pragma unroll
        for (int q = 0; q < CH; q++) {
            const int j = j_first + q;
            if (j < m) {
                rmn = min_lt(rmn, volr[q]);
                rmx = max_gt(rmx, volr[q]);
                if (j >= j_done) {
                    const float mn = min_lt(rmn, smn[q]);
                    const float mx = max_gt(rmx, smx[q]);
                    const float center = __fmul_rn(__fadd_rn(mx, mn), 0.5f);
                    const float a = avgr[q];
                    uint8_t s;
                    if (p.four_level) {
                        const double c = (double) center;
                        const float umid =
                            __double2float_rn(__dadd_rn(__dmul_rn((double) __fsub_rn(mx, center), 0.625), c));
                        const float lmid =
                            __double2float_rn(__dadd_rn(__dmul_rn((double) __fsub_rn(mn, center), 0.625), c));
                        s = a > center ? (a > umid ? 1 : 0) : (a < lmid ? 3 : 2);
                    } else {
                        s = a > center ? (p.invert ? 0 : 1) : (p.invert ? 1 : 0);
                    }
                    sym_out[q] = s;
                }
            }
        }
        emitted += m - j_done;
        if (!full) {
            j_done = m;
            break;
        }

        // next block: this block's volumes become the "previous" ring content
        P = P_next;
        vo = vo_next;
        j_done = 0;
        m = m_next;
#pragma unroll
        for (int q = 0; q < CH; q++) pv[q] = volr[q];
        suffix_of_previous<G>(pv, smn, smx, gl, gmask);
        cp_async_wait_all();
        __syncwarp(gmask);
    }

    // carry: state + the unconsumed tail [P, T), written right-aligned in front of column carry_cap of the row
    // the NEXT call reads (the two work buffers alternate so that the producer of the next chunk can already
    // run while this kernel is still reading the current one)
#pragma unroll
    for (int q = 0; q < CH; q++) {
        if (j_first + q < kBlockSyms) st->vol_prev[j_first + q] = pv[q];
    }
    const int keep = T - P;
    {
        // through the (now idle) sample buffer, so that all loads are in flight at once instead of one
        // load->store round trip per element
        float* dst = p.work_next + (size_t) ch * p.pitch + p.carry_cap - keep;
        __syncwarp(gmask);   // every lane of the group is done reading the staged block (window sums of a partial block)
        for (int idx = gl; idx < keep; idx += G) cp_async4(S + idx, at(P + idx));
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp(gmask);
        for (int idx = gl; idx < keep; idx += G) dst[idx] = S[idx];
    }
    if (gl == 0) {
        st->vo = vo;
        st->j_done = j_done;
        st->carry_len = keep;
        p.nsym[ch] = (uint32_t) emitted;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Split form of K2 (dh_demod_set_split).  The only sequential dependency of the demodulator is the +-1 nudge that the
// variance search of block b hands to block b + 1; the window sums, the volume ring and the slicing of a block depend
// on nothing but that block's position {P, nudge}.  So the chain is walked by a kernel that does nothing else
// (demod_search_kernel: same lane groups and staging as demod_kernel, a third of its instructions per block), which
// records {P, nudge} of every block; two wide kernels without any serial part finish the job:
//   demod_volume_kernel  one lane per symbol: window sums -> {volume, average} (gfsk_demodulator.cpp:28-35, 82-83, 88)
//   demod_slice_kernel   one 10-lane group per (channel, block): ring min / max = (prefix over this block's volumes) x
//                        (suffix over the previous block's) by the same shuffle scans as demod_kernel, thresholds,
//                        symbols (gfsk_demodulator.cpp:88-122)
// Arithmetic, rounding and operation order are those of demod_kernel (the helpers are shared), only the schedule
// differs.  The per-channel state alternates between two sets so that the slice group of block 0 can read the previous
// call's volume ring while the group of the last full block already writes the next one.
// ---------------------------------------------------------------------------------------------------------------------

// the logical sample stream of one channel: carried tail (work row) followed by the chunk (same row, or the caller's)
struct RowView {
    const float* row0;      // &stream[0] inside the work row
    const float* ext_row;   // &chunk[0] inside the caller's rows, or nullptr: the chunk follows the tail in the work row
    int carry_len;
    int col0;
    int T;                  // samples visible to this call
    __device__ __forceinline__ RowView(const DemodParams& p, int ch, int carry_len_) {
        carry_len = carry_len_;
        col0 = p.carry_cap - carry_len;
        row0 = p.work + (size_t) ch * p.pitch + col0;
        ext_row = p.ext ? p.ext + (size_t) ch * p.ext_pitch : nullptr;
        T = carry_len + p.n;
    }
    __device__ __forceinline__ const float* at(int x) const {
        return (ext_row != nullptr && x >= carry_len) ? ext_row + (x - carry_len) : row0 + x;
    }
    // offset of sample P inside its 16-byte unit (0 for a range that starts in the tail of an `ext` call: it is
    // copied element by element because it may straddle two buffers)
    __device__ __forceinline__ int align_of(int P) const {
        if (ext_row == nullptr) return (col0 + P) & 3;
        return P >= carry_len ? (P - carry_len) & 3 : 0;
    }
    // asynchronous staging of [P, P + len) by W cooperating lanes (this one is lane l of them) with aligned 16-byte
    // copies; sample P + x lands at S[align_of(P) + x]
    template <int W>
    __device__ __forceinline__ void stage(float* S, int P, int len, int l) const {
        const int a0 = align_of(P);
        if (ext_row == nullptr) {
            const float4* src = reinterpret_cast<const float4*>(row0 + P - a0);
            float4* dst = reinterpret_cast<float4*>(S);
            const int nvec = (a0 + len + 3) >> 2;
            for (int v = l; v < nvec; v += W) cp_async16(dst + v, src + v);
        } else if (P >= carry_len) {
            // whole vectors as long as they end inside the caller's row, single samples behind them (never read
            // past sample T - 1: the buffer is not ours)
            const float* base = ext_row + (P - carry_len) - a0;
            const int avail = T - P + a0;
            const int want = a0 + len;
            const int nvec = min(want, avail) >> 2;
            const float4* src = reinterpret_cast<const float4*>(base);
            float4* dst = reinterpret_cast<float4*>(S);
            for (int v = l; v < nvec; v += W) cp_async16(dst + v, src + v);
            for (int x = 4 * nvec + l; x < min(want, avail); x += W) cp_async4(S + x, base + x);
        } else {
            for (int x = l; x < len && P + x < T; x += W) cp_async4(S + x, at(P + x));
        }
        cp_async_commit();
    }
};

// K2a: the nudge chain.  Lane groups, staging and the search itself are those of demod_kernel; per block it records
// where the block starts and which nudge is pending, at the end it carries the unconsumed tail and the scalar state.
template <int G, int SPS, int THREADS>
__global__ void __launch_bounds__(THREADS) demod_search_kernel(const __grid_constant__ DemodParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int kPerWarp = Group<G>::kPerWarp;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp_in_warp = lane / G;
    if (grp_in_warp >= kPerWarp) return;
    const int gl = lane - grp_in_warp * G;
    const int grp = warp * kPerWarp + grp_in_warp;
    const int ch = blockIdx.x * ((THREADS / 32) * kPerWarp) + grp;
    if (ch >= p.channels) return;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (grp_in_warp * G);

    const int sps = SPS > 0 ? SPS : p.sps;
    // one or two sample buffers (p.search_dbuf), then the phase variances
    float* const Sbuf0 = smem + (size_t) grp * p.group_floats;
    float* const Sbuf1 = p.search_dbuf ? Sbuf0 + p.samples_cap : Sbuf0;
    double* var = reinterpret_cast<double*>(Sbuf0 + (p.search_dbuf ? 2 : 1) * p.samples_cap);
    float* S = Sbuf0;        // buffer that holds the current block

    const ChannelState* st = p.state + ch;
    int vo = st->vo;
    const int j_done_in = st->j_done;
    const int carry_len = st->carry_len;
    const RowView view(p, ch, carry_len);
    const int T = view.T;
    int2* rec = p.rec + (size_t) ch * p.rec_pitch;
    const int full_len = kBlockSyms * sps + 2;   // what demod_kernel stages for a complete block

    int P = 0;
    int nfull = 0;
    int m = processable(T, P, vo, sps);
    if (gl == 0) rec[0] = make_int2(P, vo);
    if (m == kBlockSyms) {
        view.stage<G>(S, P, full_len, gl);
        cp_async_wait_all();
        __syncwarp(gmask);
    }
    while (m == kBlockSyms) {
        const int a0 = view.align_of(P);
        // Where the next block starts does not depend on the outcome of this search, only whether it is complete
        // does: with two buffers its samples are requested now and arrive while this block is searched.  It can only
        // be complete if at least 100 * sps + 1 samples are visible from its start.
        const int P_next = P + kBlockSyms * sps + vo;
        float* const S_other = S == Sbuf0 ? Sbuf1 : Sbuf0;
        const bool prestaged = p.search_dbuf && T - P_next >= kBlockSyms * sps + 1;
        if (prestaged) view.stage<G>(S_other, P_next, full_len, gl);
        // variance-minimum phase search over the 100 windows of the block (gfsk_demodulator.cpp:41-80).  This kernel
        // is nothing but these two ordered chains, so on the compile-time paths they are fully unrolled: the 100
        // samples of a phase stay in registers between the two passes and their loads / conversions are scheduled
        // ahead of the dependent FADD / DADD chain (134 registers; demod_kernel keeps `#pragma unroll 11` because its
        // slicer state already fills the register file)
        for (int i = gl; i < sps; i += G) {
            const float* w0 = S + a0 + i;
            const float* wv = w0 + vo;          // windows 1..99 are shifted by the pending nudge
            float total = __fadd_rn(0.0f, w0[0]);
            if (SPS > 0) {
#pragma unroll
                for (int k = 1; k < kBlockSyms; k++) total = __fadd_rn(total, wv[k * SPS]);
            } else {
                for (int k = 1; k < kBlockSyms; k++) total = __fadd_rn(total, wv[k * sps]);
            }
            const double mean = (double) (SPS > 0 ? div_by_const(total, 1.0 / 100.0) : __fdiv_rn(total, 100.0f));
            double d = __dsub_rn(mean, (double) w0[0]);
            double dsum = __dadd_rn(0.0, __dmul_rn(d, d));
            if (SPS > 0) {
#pragma unroll
                for (int k = 1; k < kBlockSyms; k++) {
                    d = __dsub_rn(mean, (double) wv[k * SPS]);
                    dsum = __dadd_rn(dsum, __dmul_rn(d, d));
                }
            } else {
                for (int k = 1; k < kBlockSyms; k++) {
                    d = __dsub_rn(mean, (double) wv[k * sps]);
                    dsum = __dadd_rn(dsum, __dmul_rn(d, d));
                }
            }
            var[i] = __ddiv_rn(dsum, 100.0);
        }
        __syncwarp(gmask);
        double vmin = var[0];
        int vpos = 0;
        for (int i = 1; i < sps; i++) {
            const double v = var[i];
            if (v < vmin) {
                vmin = v;
                vpos = i;
            }
        }
        int vo_next = 0;
        if (vmin <= 0 || vmin > 5000000) {
            // no decision
        } else if (vpos > 0 && vpos < sps / 2) {
            vo_next = +1;
        } else if (vpos >= sps / 2 && vpos < sps - 1) {
            vo_next = -1;
        }
        const int m_next = processable(T, P_next, vo_next, sps);
        __syncwarp(gmask);   // every lane of the group is done with the staged block and with var[]
        if (m_next == kBlockSyms && !prestaged) view.stage<G>(S_other, P_next, full_len, gl);
        S = S_other;
        nfull++;
        P = P_next;
        vo = vo_next;
        m = m_next;
        if (gl == 0 && nfull < p.rec_pitch) rec[nfull] = make_int2(P, vo);
        cp_async_wait_all();
        __syncwarp(gmask);
    }

    // block `nfull` is incomplete (m < 100 symbols, possibly none): its windows are recomputed by the next call from
    // the carried tail [P, T), written right-aligned in front of column carry_cap of the rows the NEXT call reads
    const int keep = T - P;
    {
        float* dst = p.work_next + (size_t) ch * p.pitch + p.carry_cap - keep;
        __syncwarp(gmask);
        for (int idx = gl; idx < keep; idx += G) cp_async4(S + idx, view.at(P + idx));
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp(gmask);
        for (int idx = gl; idx < keep; idx += G) dst[idx] = S[idx];
    }
    if (gl == 0) {
        const int j_cur = nfull == 0 ? j_done_in : 0;   // symbols of block `nfull` emitted by earlier calls
        ChannelState* so = p.state_out + ch;
        so->vo = vo;
        so->j_done = m > j_cur ? m : j_cur;
        so->carry_len = keep;
        p.hdr[ch] = make_int4(nfull, m, j_done_in, carry_len);
        const int emitted = nfull == 0 ? (m > j_done_in ? m - j_done_in : 0)
                                       : (kBlockSyms - j_done_in) + (nfull - 1) * kBlockSyms + m;
        p.nsym[ch] = (uint32_t) emitted;
    }
}

// K2b: one lane per symbol.  Symbols are numbered b * 100 + j over the blocks of this call; a warp stages the (contiguous)
// sample range its 32 windows cover and every lane sums its own window in the reference's order.
constexpr int kVolThreads = 256;
__host__ __device__ constexpr int vol_warp_floats(int sps) { return (32 * sps + 2 + 3 + 3 + 3) & ~3; }

template <int SPS>
__global__ void __launch_bounds__(kVolThreads) demod_volume_kernel(const __grid_constant__ DemodParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ch = blockIdx.x;
    const int sps = SPS > 0 ? SPS : p.sps;
    const int4 hd = p.hdr[ch];
    const int total = hd.x * kBlockSyms + hd.y;   // symbols that have a complete window in this call
    const int s0 = (blockIdx.y * (kVolThreads / 32) + warp) * 32;
    if (s0 >= total) return;                      // warp-uniform
    const int s = s0 + lane;
    const bool active = s < total;
    const int b = s / kBlockSyms;
    const int j = s - b * kBlockSyms;
    int start = 0;
    if (active) {
        const int2 r = p.rec[(size_t) ch * p.rec_pitch + b];
        start = r.x + j * sps + (j ? r.y : 0);    // windows 1..99 of a block are shifted by its pending nudge
    }
    // window starts grow with the symbol number (also across blocks: P of block b + 1 is where window 99 of block b
    // ends), so the warp's windows cover [start of lane 0, start of the last active lane + sps)
    const int nact = min(32, total - s0);
    const int lo = __shfl_sync(0xffffffffu, start, 0);
    const int hi = __shfl_sync(0xffffffffu, start, nact - 1) + sps;
    const RowView view(p, ch, hd.w);
    float* S = smem + (size_t) warp * vol_warp_floats(sps);
    const int a0 = view.align_of(lo);
    view.stage<32>(S, lo, hi - lo, lane);
    cp_async_wait_all();
    __syncwarp();
    if (!active) return;

    const float* w = S + a0 + (start - lo);
    float sum = 0.0f, vsum = 0.0f, vol, avg;
    if (SPS > 0) {
        constexpr int kLo = eval_lo(SPS > 0 ? SPS : 10), kHi = eval_hi(SPS > 0 ? SPS : 10);
#pragma unroll
        for (int i = 0; i < SPS; i++) {
            const float v = w[i];
            if (i >= kLo && i < kHi) sum = __fadd_rn(sum, v);
            vsum = __fadd_rn(vsum, v);
        }
        vol = div_by_const(vsum, 1.0 / (double) (SPS > 0 ? SPS : 1));
        avg = kHi - kLo == 4 ? __fmul_rn(sum, 0.25f) : div_by_const(sum, 1.0 / (double) (kHi - kLo));
    } else {
        for (int i = 0; i < sps; i++) {
            const float v = w[i];
            if (i >= p.lo && i < p.hi) sum = __fadd_rn(sum, v);
            vsum = __fadd_rn(vsum, v);
        }
        vol = __fdiv_rn(vsum, (float) sps);
        avg = __fdiv_rn(sum, (float) (p.hi - p.lo));
    }
    p.va[(size_t) ch * p.rec_pitch * kBlockSyms + s] = make_float2(vol, avg);
}

// K2c: calibrateAudio + slicing of one block per 10-lane group (three groups per warp), lane gl owns the symbols
// [10 gl, 10 gl + 10) like in demod_kernel<10, ...>.
constexpr int kSliceG = 10;
__global__ void __launch_bounds__(kThreads) demod_slice_kernel(const __grid_constant__ DemodParams p) {
    constexpr int G = kSliceG;
    constexpr int kPerWarp = Group<G>::kPerWarp;
    constexpr int CH = Group<G>::kChunk;
    static_assert(CH * G == kBlockSyms && CH % 2 == 0, "the vector loads below assume 10 symbols per lane");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp_in_warp = lane / G;
    if (grp_in_warp >= kPerWarp) return;
    const int gl = lane - grp_in_warp * G;
    const int grp = warp * kPerWarp + grp_in_warp;
    const int ch = blockIdx.x * ((kThreads / 32) * kPerWarp) + grp;
    if (ch >= p.channels) return;
    const unsigned gmask = ((1u << G) - 1u) << (grp_in_warp * G);

    const int b = blockIdx.y;
    const int4 hd = p.hdr[ch];
    const int nfull = hd.x;
    if (b > nfull) return;                                  // the whole group leaves together
    const int m = b < nfull ? kBlockSyms : hd.y;            // symbols of this block that have a window
    const int j_done_in = hd.z;
    const int j_done = b == 0 ? j_done_in : 0;              // symbols of this block emitted by earlier calls
    if (m <= j_done && nfull != 0) return;                  // nothing to emit and not in charge of the ring hand-over
    const int j_first = gl * CH;

    // {volume, average} of the own symbols; entries without a window (j >= m) were not written by this call
    const float2* va = p.va + ((size_t) ch * p.rec_pitch + b) * kBlockSyms + j_first;
    float volr[CH], avgr[CH], pv[CH], smn[CH], smx[CH];
#pragma unroll
    for (int q = 0; q < CH; q += 2) {
        const float4 t = *reinterpret_cast<const float4*>(va + q);
        volr[q] = j_first + q < m ? t.x : 0.0f;
        avgr[q] = j_first + q < m ? t.y : 0.0f;
        volr[q + 1] = j_first + q + 1 < m ? t.z : 0.0f;
        avgr[q + 1] = j_first + q + 1 < m ? t.w : 0.0f;
    }
    // the ring content this block finds: the previous block's volumes (block 0: what the previous call left)
    if (b == 0) {
        const ChannelState* st = p.state + ch;
#pragma unroll
        for (int q = 0; q < CH; q++) pv[q] = st->vol_prev[j_first + q];
    } else {
#pragma unroll
        for (int q = 0; q < CH; q += 2) {
            const float4 t = *reinterpret_cast<const float4*>(va - kBlockSyms + q);
            pv[q] = t.x;
            pv[q + 1] = t.z;
        }
    }
    suffix_of_previous<G>(pv, smn, smx, gl, gmask);

    float emn = FLT_MAX, emx = FLT_MIN;
#pragma unroll
    for (int q = 0; q < CH; q++) {
        if (j_first + q < m) {
            emn = min_lt(emn, volr[q]);
            emx = max_gt(emx, volr[q]);
        }
    }
    exclusive_up<G>(gmask, gl, emn, emx);

    // calibrateAudio + slicing (gfsk_demodulator.cpp:88-104, 109-122); symbol j of block b is output number
    // b * 100 + j - j_done_in of this call
    float rmn = emn, rmx = emx;
    uint8_t* const sym_out = p.sym + (size_t) ch * p.sym_pitch + (b * kBlockSyms - j_done_in + j_first);
#pragma unroll
    for (int q = 0; q < CH; q++) {
        const int j = j_first + q;
        if (j < m) {
            rmn = min_lt(rmn, volr[q]);
            rmx = max_gt(rmx, volr[q]);
            if (j >= j_done) {
                const float mn = min_lt(rmn, smn[q]);
                const float mx = max_gt(rmx, smx[q]);
                const float center = __fmul_rn(__fadd_rn(mx, mn), 0.5f);
                const float a = avgr[q];
                uint8_t s;
                if (p.four_level) {
                    const double c = (double) center;
                    const float umid =
                        __double2float_rn(__dadd_rn(__dmul_rn((double) __fsub_rn(mx, center), 0.625), c));
                    const float lmid =
                        __double2float_rn(__dadd_rn(__dmul_rn((double) __fsub_rn(mn, center), 0.625), c));
                    s = a > center ? (a > umid ? 1 : 0) : (a < lmid ? 3 : 2);
                } else {
                    s = a > center ? (p.invert ? 0 : 1) : (p.invert ? 1 : 0);
                }
                sym_out[q] = s;
            }
        }
    }

    // ring hand-over to the next call: the volumes of the last complete block, or the unchanged ring without one
    if (b == nfull - 1) {
        ChannelState* so = p.state_out + ch;
#pragma unroll
        for (int q = 0; q < CH; q++) so->vol_prev[j_first + q] = volr[q];
    } else if (nfull == 0) {
        ChannelState* so = p.state_out + ch;
#pragma unroll
        for (int q = 0; q < CH; q++) so->vol_prev[j_first + q] = pv[q];
    }
}

}  // namespace

struct dh_demod {
    int device = 0;
    uint32_t channels = 0;
    int four_level = 0;
    int sps = 0;
    int invert = 0;
    int lo = 0, hi = 0;
    int carry_cap = 0;
    ChannelState* d_state = nullptr;
    float* d_work[2] = {nullptr, nullptr};   // alternate per call (see demod_kernel carry)
    int cur = 0;
    size_t pitch = 0;      // elements per work row
    size_t max_n = 0;      // chunk capacity of the work rows
    bool smem_attr_set = false;   // the kernel variant and its dynamic smem size are fixed per bank
    // split mode (dh_demod_set_split): second state set, per-call block records and per-symbol {volume, average}
    int split = 0;                 // -1 auto, 0 one kernel, 1 three kernels
    bool last_split = false;       // schedule of the most recent process call
    ChannelState* d_state_set[2] = {nullptr, nullptr};   // d_state_set[0] == the set allocated at create
    int scur = 0;                                        // d_state == d_state_set[scur]
    int2* d_rec = nullptr;
    int4* d_hdr = nullptr;
    float2* d_va = nullptr;
    size_t rec_pitch = 0;
    bool split_attr_set = false;
};

namespace {

int demod_reserve(dh_demod* h, size_t max_n) {
    if (h->d_work[0] && max_n <= h->max_n) return DH_OK;
    const size_t n4 = (max_n + 3) & ~(size_t) 3;
    const size_t pitch = (size_t) h->carry_cap + n4;
    float* nw[2] = {nullptr, nullptr};
    // + one 16-byte vector: the staging copies of the kernel are whole float4s and the range they cover may end one
    // sample past the last row (a block staged with a pending -1 nudge)
    const size_t alloc_bytes = ((size_t) h->channels * pitch + 4) * sizeof(float);
    for (int b = 0; b < 2; b++) {
        DH_CUDA(cudaMalloc(&nw[b], alloc_bytes));
        DH_CUDA(cudaMemset(nw[b], 0, alloc_bytes));
    }
    if (h->d_work[0]) {
        // keep the carried tails (they live in the buffer the next call reads); the bank may be mid-stream
        DH_CUDA(cudaDeviceSynchronize());
        DH_CUDA(cudaMemcpy2D(nw[h->cur], pitch * sizeof(float), h->d_work[h->cur], h->pitch * sizeof(float),
                             (size_t) h->carry_cap * sizeof(float), h->channels, cudaMemcpyDeviceToDevice));
        DH_CUDA(cudaFree(h->d_work[0]));
        DH_CUDA(cudaFree(h->d_work[1]));
    }
    h->d_work[0] = nw[0];
    h->d_work[1] = nw[1];
    h->pitch = pitch;
    h->max_n = n4;
    return DH_OK;
}

// split mode: blocks one call of n samples can touch per channel, counting the trailing partial one (every complete
// block consumes at least 100 * sps - 1 samples of the at most carry_cap + n visible ones)
size_t split_blocks(const dh_demod* h, size_t n) {
    return ((size_t) h->carry_cap + n) / ((size_t) kBlockSyms * h->sps - 1) + 1;
}

// second state set + per-call scratch of the split kernels, grown on demand (scratch only: nothing to preserve)
int split_reserve(dh_demod* h, size_t n) {
    const size_t ch = h->channels;
    if (!h->d_state_set[1]) {
        DH_CUDA(cudaMalloc(&h->d_state_set[1], ch * sizeof(ChannelState)));
        DH_CUDA(cudaMemset(h->d_state_set[1], 0, ch * sizeof(ChannelState)));
    }
    if (!h->d_hdr) DH_CUDA(cudaMalloc(&h->d_hdr, ch * sizeof(int4)));
    const size_t need = split_blocks(h, n) + 1;
    if (need > h->rec_pitch) {
        DH_CUDA(cudaDeviceSynchronize());
        cudaFree(h->d_rec);
        cudaFree(h->d_va);
        h->d_rec = nullptr;
        h->d_va = nullptr;
        h->rec_pitch = 0;
        DH_CUDA(cudaMalloc(&h->d_rec, ch * need * sizeof(int2)));
        DH_CUDA(cudaMalloc(&h->d_va, ch * need * kBlockSyms * sizeof(float2)));
        DH_CUDA(cudaMemset(h->d_va, 0, ch * need * kBlockSyms * sizeof(float2)));
        h->rec_pitch = need;
    }
    return DH_OK;
}

}  // namespace

extern "C" {

int dh_demod_create(dh_demod** out, int device, uint32_t channels, int four_level, uint32_t sps, int invert) {
    DH_REQUIRE(out != nullptr, DH_E_INVALID, "dh_demod_create: out is NULL");
    *out = nullptr;
    DH_REQUIRE(channels > 0, DH_E_INVALID, "dh_demod_create: channels must be > 0");
    DH_REQUIRE(sps >= 4 && sps <= 128, DH_E_UNSUPPORTED, "dh_demod_create: samplesPerSymbol=%u not supported (4..128)",
               sps);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        dh::set_error("dh_demod_create: no CUDA device available (this library has no CPU fallback)");
        return DH_E_NODEVICE;
    }
    DH_REQUIRE(device >= 0 && device < ndev, DH_E_INVALID, "dh_demod_create: device %d out of range", device);
    dh::DeviceGuard guard(device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", device);
    dh_demod* h = new (std::nothrow) dh_demod();
    DH_REQUIRE(h != nullptr, DH_E_NOMEM, "dh_demod_create: out of host memory");
    h->device = device;
    h->channels = channels;
    h->four_level = four_level != 0;
    h->sps = (int) sps;
    h->invert = invert != 0;
    // lowestEval / highestEval (gfsk_demodulator.cpp:8-9): roundf of a float quotient
    h->lo = (int) roundf((float) sps / 3);
    h->hi = (int) roundf((float) sps * 2 / 3);
    h->carry_cap = kBlockSyms * (int) sps + kCarrySlack;
    cudaError_t e = cudaMalloc(&h->d_state, (size_t) channels * sizeof(ChannelState));
    if (e == cudaSuccess) e = cudaMemset(h->d_state, 0, (size_t) channels * sizeof(ChannelState));
    if (e != cudaSuccess) {
        dh::set_error("dh_demod_create: %s", cudaGetErrorString(e));
        cudaFree(h->d_state);
        delete h;
        return (int) e;
    }
    h->d_state_set[0] = h->d_state;
    static const int split_default = getenv("DH_DEMOD_SPLIT") ? atoi(getenv("DH_DEMOD_SPLIT")) : kSplitDefault;
    h->split = split_default < 0 ? -1 : (split_default != 0);
    *out = h;
    return DH_OK;
}

int dh_demod_set_split(dh_demod* h, int enable) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_demod_set_split: handle is NULL");
    h->split = enable < 0 ? -1 : (enable != 0);
    return DH_OK;
}

int dh_demod_kernels_per_call(const dh_demod* h) {
    if (!h) return 0;
    if (h->split >= 0) return h->split ? 3 : 1;
    return h->last_split ? 3 : 1;
}

int dh_demod_reserve(dh_demod* h, size_t max_n, float** d_buf, size_t* pitch) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_demod_reserve: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    int rc = demod_reserve(h, max_n);
    if (rc != DH_OK) return rc;
    if (d_buf) *d_buf = h->d_work[h->cur] + h->carry_cap;
    if (pitch) *pitch = h->pitch;
    return DH_OK;
}

size_t dh_demod_max_symbols(const dh_demod* h, size_t n) {
    if (!h) return 0;
    return ((size_t) h->carry_cap + n) / (size_t) (h->sps - 1) + 2;
}

int dh_demod_process(dh_demod* h, const float* d_in, size_t in_pitch, size_t n, uint8_t* d_sym, size_t sym_pitch,
                     uint32_t* d_nsym, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_demod_process: handle is NULL");
    DH_REQUIRE(d_sym != nullptr && d_nsym != nullptr, DH_E_INVALID, "dh_demod_process: NULL output buffer");
    DH_REQUIRE(n <= 0x40000000u, DH_E_INVALID, "dh_demod_process: n too large");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    cudaStream_t st = (cudaStream_t) stream;
    if (n == 0) {
        DH_CUDA(cudaMemsetAsync(d_nsym, 0, (size_t) h->channels * sizeof(uint32_t), st));
        return DH_OK;
    }
    DH_REQUIRE(d_in != nullptr, DH_E_INVALID, "dh_demod_process: NULL input buffer");
    DH_REQUIRE(sym_pitch >= dh_demod_max_symbols(h, n), DH_E_INVALID,
               "dh_demod_process: sym_pitch %zu too small, need dh_demod_max_symbols(n) = %zu", sym_pitch,
               dh_demod_max_symbols(h, n));
    const bool zero_copy =
        h->d_work[0] && d_in == h->d_work[h->cur] + h->carry_cap && in_pitch == h->pitch && n <= h->max_n;
    // any other 16-byte aligned device buffer is read in place (only the carried tails live in the work rows);
    // unaligned rows are copied behind the tails first
    const bool in_place = !zero_copy && reinterpret_cast<uintptr_t>(d_in) % 16 == 0 && in_pitch % 4 == 0;
    if (!zero_copy) {
        DH_REQUIRE(in_pitch >= n, DH_E_INVALID, "dh_demod_process: in_pitch < n");
        int rc = demod_reserve(h, in_place ? 0 : n);
        if (rc != DH_OK) return rc;
        if (!in_place) {
            DH_CUDA(cudaMemcpy2DAsync(h->d_work[h->cur] + h->carry_cap, h->pitch * sizeof(float), d_in,
                                      in_pitch * sizeof(float), n * sizeof(float), h->channels, cudaMemcpyDeviceToDevice, st));
        }
    }

    DemodParams p;
    p.work = h->d_work[h->cur];
    p.work_next = h->d_work[h->cur ^ 1];
    p.ext = in_place ? d_in : nullptr;
    p.ext_pitch = in_pitch;
    p.pitch = h->pitch;
    p.sym = d_sym;
    p.sym_pitch = sym_pitch;
    p.nsym = d_nsym;
    p.state = h->d_state;
    p.channels = (int) h->channels;
    p.n = (int) n;
    p.sps = h->sps;
    p.lo = h->lo;
    p.hi = h->hi;
    p.four_level = h->four_level;
    p.invert = h->invert;
    p.carry_cap = h->carry_cap;
    p.samples_cap = (kBlockSyms * h->sps + 2 + 3 + 3 + 3) & ~3;
    // samples | var (doubles; 8-byte aligned because samples_cap is a multiple of 4)
    p.group_floats = p.samples_cap + 2 * ((h->sps + 1) & ~1);
    // Three 10-lane groups of a warp read 10 consecutive floats each in the variance search: with group segments
    // that are 12 banks apart (mod 32) their bank ranges do not overlap (even, so the doubles stay 8-byte aligned)
    if (h->sps == 10 || h->sps == 20) p.group_floats += (12 - p.group_floats % 32 + 32) % 32;

    // lanes per channel: 10 on the compile-time fast paths (3 channels per warp), else 16 or 32
    const bool fast = (h->sps == 10 || h->sps == 20 || h->sps == 40) && h->lo == eval_lo(h->sps) && h->hi == eval_hi(h->sps);
    // sps = 40: the staged block (16 KB per channel) caps the channels in flight per SM whatever the group width, so
    // the lane group is widened to 20 (two phase passes instead of four, 5 symbols per lane instead of 10): the
    // latency of a block halves: 3.99 -> 3.50 ms per 32768-channel step (a full warp per channel: 7.6 ms).
    // DH_DEMOD_G40 = 10 selects the narrow variant (experiments).
    static const int g40 = getenv("DH_DEMOD_G40") ? atoi(getenv("DH_DEMOD_G40")) : 20;
    // sps = 20 likewise: 20-lane groups (one variance pass, vector window loads): NXDN pipe 2.43 -> 2.30 ms per step.
    static const int g20 = getenv("DH_DEMOD_G20") ? atoi(getenv("DH_DEMOD_G20")) : 20;
    const int G = fast ? ((h->sps == 40 && g40 != 10) || (h->sps == 20 && g20 == 20) ? 20 : 10) : (h->sps <= 16 ? 16 : 32);
    const int threads = fast && h->sps == 40 && G == 10 ? 32 : kThreads;
    const int groups = (threads / 32) * (32 / G);
    const unsigned grid = (h->channels + groups - 1) / groups;
    const size_t smem = (size_t) groups * p.group_floats * sizeof(float);
#define DH_LAUNCH_DEMOD4(GG, SS, TT, MB)                                                                             \
    do {                                                                                                             \
        if (!h->smem_attr_set)                                                                                       \
            DH_CUDA(cudaFuncSetAttribute(demod_kernel<GG, SS, TT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         (int) smem));                                                               \
        demod_kernel<GG, SS, TT, MB><<<grid, TT, smem, st>>>(p);                                                      \
    } while (0)
#define DH_LAUNCH_DEMOD(GG, SS, TT)                                                                                  \
    do {                                                                                                             \
        if (!h->smem_attr_set)                                                                                       \
            DH_CUDA(cudaFuncSetAttribute(demod_kernel<GG, SS, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                         (int) smem));                                                               \
        demod_kernel<GG, SS, TT><<<grid, TT, smem, st>>>(p);                                                          \
    } while (0)
    // ---- split mode: search chain -> per-symbol window sums -> per-block slicing (three kernels, same stream) ----
    const size_t nblk = split_blocks(h, n);                                      // blocks incl. the partial one
    const size_t vol_tiles = (nblk * kBlockSyms + kVolThreads - 1) / kVolThreads;
    const bool want_split = h->split > 0 || (h->split < 0 && h->channels <= kSplitAutoChannels && nblk >= kSplitAutoBlocks);
    h->last_split = false;
    if (want_split && nblk <= 65535 && vol_tiles <= 65535) {
        h->last_split = true;
        int rc = split_reserve(h, n);
        if (rc != DH_OK) return rc;
        p.state = h->d_state_set[h->scur];
        p.state_out = h->d_state_set[h->scur ^ 1];
        p.rec = h->d_rec;
        p.hdr = h->d_hdr;
        p.va = h->d_va;
        p.rec_pitch = (int) h->rec_pitch;
        // two sample buffers per group when they fit comfortably (always on the compile-time paths)
        const int group_floats_mono = p.group_floats;
        int group_floats_dbuf = p.group_floats + p.samples_cap;
        // keep the group segments 12 banks apart (see above) with the second buffer in between
        if (h->sps == 10 || h->sps == 20) group_floats_dbuf += (12 - group_floats_dbuf % 32 + 32) % 32;
        const size_t smem_dbuf = (size_t) groups * group_floats_dbuf * sizeof(float);
        p.search_dbuf = smem_dbuf <= 160 * 1024;
        const size_t smem_search = p.search_dbuf ? smem_dbuf : smem;
        if (p.search_dbuf) p.group_floats = group_floats_dbuf;
#define DH_LAUNCH_SEARCH(GG, SS, TT)                                                                                 \
    do {                                                                                                             \
        if (!h->split_attr_set)                                                                                      \
            DH_CUDA(cudaFuncSetAttribute(demod_search_kernel<GG, SS, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int) smem_search));                                                        \
        demod_search_kernel<GG, SS, TT><<<grid, TT, smem_search, st>>>(p);                                            \
    } while (0)
        if (G == 10 && h->sps == 10) {
            DH_LAUNCH_SEARCH(10, 10, kThreads);
        } else if (G == 10 && h->sps == 20) {
            DH_LAUNCH_SEARCH(10, 20, kThreads);
        } else if (G == 20 && h->sps == 20) {
            DH_LAUNCH_SEARCH(20, 20, kThreads);
        } else if (G == 20) {
            DH_LAUNCH_SEARCH(20, 40, kThreads);
        } else if (G == 10) {
            DH_LAUNCH_SEARCH(10, 40, 32);
        } else if (G == 16) {
            DH_LAUNCH_SEARCH(16, 0, kThreads);
        } else {
            DH_LAUNCH_SEARCH(32, 0, kThreads);
        }
#undef DH_LAUNCH_SEARCH
        DH_CUDA(cudaGetLastError());
        p.group_floats = group_floats_mono;
        const dim3 vgrid(h->channels, (unsigned) vol_tiles);
        const size_t vsmem = (size_t) (kVolThreads / 32) * vol_warp_floats(h->sps) * sizeof(float);
#define DH_LAUNCH_VOLUME(SS)                                                                                         \
    do {                                                                                                             \
        if (!h->split_attr_set)                                                                                      \
            DH_CUDA(cudaFuncSetAttribute(demod_volume_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) vsmem));                                                              \
        demod_volume_kernel<SS><<<vgrid, kVolThreads, vsmem, st>>>(p);                                                \
    } while (0)
        if (fast && h->sps == 10) {
            DH_LAUNCH_VOLUME(10);
        } else if (fast && h->sps == 20) {
            DH_LAUNCH_VOLUME(20);
        } else if (fast) {
            DH_LAUNCH_VOLUME(40);
        } else {
            DH_LAUNCH_VOLUME(0);
        }
#undef DH_LAUNCH_VOLUME
        DH_CUDA(cudaGetLastError());
        const unsigned sgroups = (kThreads / 32) * (32 / kSliceG);
        const dim3 sgrid((h->channels + sgroups - 1) / sgroups, (unsigned) nblk);
        demod_slice_kernel<<<sgrid, kThreads, 0, st>>>(p);
        DH_CUDA(cudaGetLastError());
        h->split_attr_set = true;
        h->scur ^= 1;
        h->d_state = h->d_state_set[h->scur];   // reset / export / import address the current set
        h->cur ^= 1;
        return DH_OK;
    }

    static const int minb = getenv("DH_DEMOD_MINB") ? atoi(getenv("DH_DEMOD_MINB")) : 0;   // experiment switch
    if (G == 10 && h->sps == 10 && minb == 10) {
        DH_LAUNCH_DEMOD4(10, 10, kThreads, 10);
    } else if (G == 10 && h->sps == 10 && minb == 12) {
        DH_LAUNCH_DEMOD4(10, 10, kThreads, 12);
    } else if (G == 10 && h->sps == 10 && minb == 16) {
        DH_LAUNCH_DEMOD4(10, 10, kThreads, 16);
    } else if (G == 10 && h->sps == 10) {
        DH_LAUNCH_DEMOD(10, 10, kThreads);
    } else if (G == 10 && h->sps == 20) {
        DH_LAUNCH_DEMOD(10, 20, kThreads);
    } else if (G == 20 && h->sps == 20) {
        DH_LAUNCH_DEMOD(20, 20, kThreads);
    } else if (G == 20) {
        DH_LAUNCH_DEMOD(20, 40, kThreads);

    } else if (G == 10) {
        DH_LAUNCH_DEMOD(10, 40, 32);   // 16 KB of staged samples per channel: one warp (3 channels) per CTA
    } else if (G == 16) {
        DH_LAUNCH_DEMOD(16, 0, kThreads);
    } else {
        DH_LAUNCH_DEMOD(32, 0, kThreads);
    }
#undef DH_LAUNCH_DEMOD
#undef DH_LAUNCH_DEMOD4
    DH_CUDA(cudaGetLastError());
    h->smem_attr_set = true;
    h->cur ^= 1;   // the carried tails now sit in the other buffer
    return DH_OK;
}

int dh_demod_reset(dh_demod* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_demod_reset: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    DH_CUDA(cudaMemsetAsync(h->d_state, 0, (size_t) h->channels * sizeof(ChannelState), (cudaStream_t) stream));
    return DH_OK;
}

uint32_t dh_demod_channels(const dh_demod* h) { return h ? h->channels : 0; }

// ---- state: ChannelState of every channel + the unconsumed sample tails (columns [0, carry_cap) of the work rows) --
static dh::StateHeader demod_header(const dh_demod* h) {
    const uint64_t payload = (uint64_t) h->channels * (sizeof(ChannelState) + (size_t) h->carry_cap * sizeof(float));
    return dh::make_state_header(2, h->channels, (uint32_t) h->sps, (uint32_t) h->four_level, (uint32_t) h->invert,
                                 (uint32_t) h->carry_cap, payload);
}

int dh_demod_state_size(const dh_demod* h, size_t* bytes) {
    DH_REQUIRE(h != nullptr && bytes != nullptr, DH_E_INVALID, "dh_demod_state_size: NULL argument");
    *bytes = sizeof(dh::StateHeader) + demod_header(h).payload;
    return DH_OK;
}

int dh_demod_state_export(dh_demod* h, void* h_buf, size_t cap, size_t* written, void* stream) {
    DH_REQUIRE(h != nullptr && h_buf != nullptr, DH_E_INVALID, "dh_demod_state_export: NULL argument");
    const dh::StateHeader hd = demod_header(h);
    DH_REQUIRE(cap >= sizeof(hd) + hd.payload, DH_E_INVALID, "dh_demod_state_export: buffer too small");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    int rc = demod_reserve(h, 0);
    if (rc != DH_OK) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    char* out = static_cast<char*>(h_buf);
    std::memcpy(out, &hd, sizeof(hd));
    out += sizeof(hd);
    DH_CUDA(cudaMemcpyAsync(out, h->d_state, (size_t) h->channels * sizeof(ChannelState), cudaMemcpyDeviceToHost, st));
    out += (size_t) h->channels * sizeof(ChannelState);
    const size_t w = (size_t) h->carry_cap * sizeof(float);
    DH_CUDA(cudaMemcpy2DAsync(out, w, h->d_work[h->cur], h->pitch * sizeof(float), w, h->channels, cudaMemcpyDeviceToHost,
                              st));
    DH_CUDA(cudaStreamSynchronize(st));
    if (written) *written = sizeof(hd) + hd.payload;
    return DH_OK;
}

int dh_demod_state_import(dh_demod* h, const void* h_buf, size_t bytes, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_demod_state_import: handle is NULL");
    const dh::StateHeader hd = demod_header(h);
    int rc = dh::check_state_header(h_buf, bytes, hd, "dh_demod_state_import");
    if (rc != DH_OK) return rc;
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    rc = demod_reserve(h, 0);
    if (rc != DH_OK) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    const char* in = static_cast<const char*>(h_buf) + sizeof(hd);
    DH_CUDA(cudaMemcpyAsync(h->d_state, in, (size_t) h->channels * sizeof(ChannelState), cudaMemcpyHostToDevice, st));
    in += (size_t) h->channels * sizeof(ChannelState);
    const size_t w = (size_t) h->carry_cap * sizeof(float);
    DH_CUDA(cudaMemcpy2DAsync(h->d_work[h->cur], h->pitch * sizeof(float), in, w, w, h->channels, cudaMemcpyHostToDevice,
                              st));
    DH_CUDA(cudaStreamSynchronize(st));
    return DH_OK;
}

void dh_demod_destroy(dh_demod* h) {
    if (!h) return;
    dh::DeviceGuard guard(h->device);
    cudaFree(h->d_state_set[0]);
    cudaFree(h->d_state_set[1]);
    cudaFree(h->d_rec);
    cudaFree(h->d_hdr);
    cudaFree(h->d_va);
    cudaFree(h->d_work[0]);
    cudaFree(h->d_work[1]);
    delete h;
}

}  // extern "C"
