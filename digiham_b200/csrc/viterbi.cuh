// viterbi.cuh — K5: rate-1/2 K=5 hard-decision register-exchange Viterbi decoder shared by the YSF and NXDN
// decoder banks (sm_100a).  One trellis state per lane (16 states; both half-warps carry the same decode): per
// step each lane fetches the two predecessor metrics with __shfl_sync, compare-selects (ties -> predecessor k = 0,
// exactly like the references' strict `<`), then pulls the selected survivor — the whole decoded bit string
// travels with the state — word by word through __shfl_sync.
//
//   YSF  : decode_trellis, reference src/ysf_decoder/trellis.c:32-109 — metrics are uint8_t and wrap modulo 256
//          (trellis.c:28,68), every state may start the path.
//   NXDN : Digiham::Nxdn::Trellis::decode, reference src/nxdn_decoder/trellis.cpp:29-101 — uint16_t metrics (no
//          wrap within 96 steps), and the "four leading zeros" prior: during the first four steps a state whose
//          index intersects the `blocked` mask (0b1111 << step) only considers predecessor k = 0
//          (trellis.cpp:34-36,60-61,92-93).
// Both use the same transition table (trellis.c:8-25 == trellis.cpp:10-27): the expected dibit of the transition
// prev -> (outbit, prev >> 1) is linear in the bits of prev.
#pragma once
#include <stdint.h>

namespace dh {

// dibits[] (shared memory) holds STEPS received dibits; the decoded bit string comes back MSB-first in
// out_words[0 .. (STEPS+31)/32) (identical in every lane).  Returns the winning path metric.
template <int STEPS, bool NXDN>
__device__ __forceinline__ uint32_t viterbi(const uint8_t* dibits, int lane, uint32_t* out_words) {
    constexpr int NW = (STEPS + 31) / 32;
    const int state = lane & 15;
    const uint32_t outbit = (uint32_t) (state >> 3) & 1u;
    const int p0 = (state << 1) & 14;   // predecessor with k = 0; k = 1 is p0 | 1
    auto expected = [](int prev, uint32_t ob) -> uint32_t {
        uint32_t t = ob ? 3u : 0u;
        if (prev & 1) t ^= 3u;
        if (prev & 2) t ^= 2u;
        if (prev & 4) t ^= 1u;
        if (prev & 8) t ^= 1u;
        return t;
    };
    const uint32_t e0 = expected(p0, outbit), e1 = expected(p0 | 1, outbit);
    constexpr uint32_t kMask = NXDN ? 0xFFFFu : 0xFFu;
    uint32_t metric = 0;
    uint32_t surv[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) surv[w] = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int lim = (STEPS - w * 32) < 32 ? (STEPS - w * 32) : 32;
        for (int b = 0; b < lim; b++) {
            const uint32_t in = dibits[w * 32 + b] & 3u;
            const uint32_t m0 = (__shfl_sync(0xffffffffu, metric, p0) + __popc(in ^ e0)) & kMask;
            const uint32_t m1 = (__shfl_sync(0xffffffffu, metric, p0 | 1) + __popc(in ^ e1)) & kMask;
            bool take1 = m1 < m0;
            if (NXDN && w == 0 && b < 4) take1 = take1 && ((state & ((0xF << b) & 0xF)) == 0);
            const int sel = take1 ? (p0 | 1) : p0;
            metric = take1 ? m1 : m0;
#pragma unroll
            for (int v = 0; v <= w; v++) surv[v] = __shfl_sync(0xffffffffu, surv[v], sel);
            surv[w] |= outbit << (31 - b);
        }
    }
    // best = lowest state index with the minimal metric (trellis.c:94-98, trellis.cpp:86-90)
    uint32_t key = (metric << 4) | (uint32_t) state;
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, d));
    const int best = (int) (key & 15u);
#pragma unroll
    for (int w = 0; w < NW; w++) out_words[w] = __shfl_sync(0xffffffffu, surv[w], best);
    return key >> 4;
}

// Two independent decodes at once, one per half-warp (lanes 0-15 decode dibits_a, lanes 16-31 decode dibits_b):
// the single-decode form above runs the same trellis twice, this one fills the second half with useful work
// (YSF: FICH + data channel of a V/D2 frame, the two data channels of a header).  Results are returned to every
// lane: out_a / out_b and the two winning metrics.
template <int STEPS, bool NXDN>
__device__ __forceinline__ void viterbi_pair(const uint8_t* dibits_a, const uint8_t* dibits_b, int lane, uint32_t* out_a,
                                             uint32_t* out_b, uint32_t& metric_a, uint32_t& metric_b) {
    constexpr int NW = (STEPS + 31) / 32;
    const int state = lane & 15;
    const int base = lane & 16;
    const uint8_t* dibits = base ? dibits_b : dibits_a;
    const uint32_t outbit = (uint32_t) (state >> 3) & 1u;
    const int p0 = base | ((state << 1) & 14);
    auto expected = [](int prev, uint32_t ob) -> uint32_t {
        uint32_t t = ob ? 3u : 0u;
        if (prev & 1) t ^= 3u;
        if (prev & 2) t ^= 2u;
        if (prev & 4) t ^= 1u;
        if (prev & 8) t ^= 1u;
        return t;
    };
    const uint32_t e0 = expected(p0 & 15, outbit), e1 = expected((p0 & 15) | 1, outbit);
    constexpr uint32_t kMask = NXDN ? 0xFFFFu : 0xFFu;
    uint32_t metric = 0;
    uint32_t surv[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) surv[w] = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int lim = (STEPS - w * 32) < 32 ? (STEPS - w * 32) : 32;
        for (int b = 0; b < lim; b++) {
            const uint32_t in = dibits[w * 32 + b] & 3u;
            const uint32_t m0 = (__shfl_sync(0xffffffffu, metric, p0) + __popc(in ^ e0)) & kMask;
            const uint32_t m1 = (__shfl_sync(0xffffffffu, metric, p0 | 1) + __popc(in ^ e1)) & kMask;
            bool take1 = m1 < m0;
            if (NXDN && w == 0 && b < 4) take1 = take1 && ((state & ((0xF << b) & 0xF)) == 0);
            const int sel = take1 ? (p0 | 1) : p0;
            metric = take1 ? m1 : m0;
#pragma unroll
            for (int v = 0; v <= w; v++) surv[v] = __shfl_sync(0xffffffffu, surv[v], sel);
            surv[w] |= outbit << (31 - b);
        }
    }
    uint32_t key = (metric << 4) | (uint32_t) state;
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, d));   // stays inside the half
    const int best = base | (int) (key & 15u);
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const uint32_t mine = __shfl_sync(0xffffffffu, surv[w], best);
        out_a[w] = __shfl_sync(0xffffffffu, mine, 0);
        out_b[w] = __shfl_sync(0xffffffffu, mine, 16);
    }
    metric_a = __shfl_sync(0xffffffffu, key, 0) >> 4;
    metric_b = __shfl_sync(0xffffffffu, key, 16) >> 4;
}

}  // namespace dh
