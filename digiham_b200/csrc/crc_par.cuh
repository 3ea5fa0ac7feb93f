// crc_par.cuh — warp-parallel evaluation of the bit-serial CRCs of the YSF and NXDN decoders (sm_100a).
//
// Every CRC of the reference is a GF(2)-affine function of the message for a fixed length: crc(m) = C ^ XOR over
// the set message bits i of T[i].  C (the CRC of the all-zero message, final inversion included) and T[i] (the CRC
// of the unit message e_i, XOR C) are produced at compile time by running the reference's own bit-serial update
// on those messages, so the tables restate exactly
//   YSF  crc16_checksum          reference src/ysf_decoder/crc16.c:3-19      (poly 0x1021, init 0, inverted)
//   NXDN Sacch::check_crc        reference src/nxdn_decoder/sacch.cpp:76-90  (6 bit, init 0x3F, 26 message bits)
//   NXDN Facch1::check_crc       reference src/nxdn_decoder/facch1.cpp:60-75 (12 bit, init 0xFFF, 80 message bits)
// At run time lane l looks at message bits l, l + 32, ... and the partial XORs are folded with five shuffles:
// ~20 instructions per check instead of ~7 per message bit on every lane.
#pragma once
#include <stdint.h>

namespace dh {

template <int N>
struct CrcTable {
    uint16_t t[N];
    uint16_t c;
};

__host__ __device__ constexpr uint32_t crc_step_ysf16(uint32_t crc, uint32_t bit) {
    const uint32_t nx = bit ^ ((crc >> 15) & 1u);
    crc = (crc << 1) & 0xFFFFu;
    return crc ^ ((nx << 12) | (nx << 5) | nx);
}
__host__ __device__ constexpr uint32_t crc_step_nxdn6(uint32_t crc, uint32_t bit) {
    const uint32_t cb = ((crc >> 5) & 1u) ^ bit;
    if (cb) crc ^= 0x13u;
    return ((crc << 1) & 0x3Eu) | cb;
}
__host__ __device__ constexpr uint32_t crc_step_nxdn12(uint32_t crc, uint32_t bit) {
    const uint32_t cb = ((crc >> 11) & 1u) ^ bit;
    if (cb) crc ^= 0x407u;
    return ((crc << 1) & 0xFFEu) | cb;
}

// KIND: 0 = YSF crc16, 1 = NXDN crc6, 2 = NXDN crc12
template <int KIND>
__host__ __device__ constexpr uint32_t crc_of_unit(int n, int one_at) {
    uint32_t crc = KIND == 0 ? 0u : (KIND == 1 ? 0x3Fu : 0xFFFu);
    for (int i = 0; i < n; i++) {
        const uint32_t bit = i == one_at ? 1u : 0u;
        crc = KIND == 0 ? crc_step_ysf16(crc, bit) : (KIND == 1 ? crc_step_nxdn6(crc, bit) : crc_step_nxdn12(crc, bit));
    }
    return KIND == 0 ? crc ^ 0xFFFFu : crc;
}

template <int KIND, int N>
__host__ __device__ constexpr CrcTable<N> make_crc_table() {
    CrcTable<N> r = {};
    r.c = (uint16_t) crc_of_unit<KIND>(N, -1);
    for (int i = 0; i < N; i++) r.t[i] = (uint16_t) (crc_of_unit<KIND>(N, i) ^ r.c);
    return r;
}

#ifdef __CUDACC__
// words: the message, MSB first, identical in every lane; tab: the table in SHARED memory (per-lane indices)
template <int N>
__device__ __forceinline__ uint32_t crc_parallel(const uint32_t* words, const uint16_t* tab, uint32_t c0, int lane) {
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < (N + 31) / 32; k++) {
        const int i = lane + 32 * k;
        if (i < N && ((words[k] >> (31 - lane)) & 1u)) acc ^= tab[i];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, d);
    return acc ^ c0;
}
#endif

}  // namespace dh
