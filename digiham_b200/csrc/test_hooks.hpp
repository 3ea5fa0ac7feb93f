// test_hooks.hpp — launchers of the device-level FEC test kernels (dh_test_* in testapi.cu).
//
// The kernels live next to the device functions they exercise (dmr.cu, ysf.cu, pocsag.cu, nxdn.cu) and call exactly
// the functions the decoder kernels call; nothing on the product path uses them.
#pragma once
#include "common.cuh"

namespace dh {
namespace test {

// codes 0..5: hamming_7_4, hamming_13_9, hamming_15_11, hamming_16_11, quadratic_residue (16,7), golay_20_8
int dmr_fec(int code, uint32_t* d_words, uint8_t* d_ok, uint32_t n, cudaStream_t st);
// payload: [n][25] bytes (196 bits MSB first, as FramePhase packs them, dmr_phase.cpp:259-273); out: [n][12]
int dmr_bptc(const uint8_t* d_payload, uint8_t* d_out, uint8_t* d_ok, uint32_t n, cudaStream_t st);
int ysf_golay24(uint32_t* d_words, uint8_t* d_ok, uint32_t n, cudaStream_t st);
// dibits: [n][steps] one received dibit per byte; words: [n][(steps + 31) / 32] decoded bits MSB first; steps 100 / 180
int ysf_viterbi(int steps, const uint8_t* d_dibits, uint32_t n, uint32_t* d_words, uint32_t* d_metric, cudaStream_t st);
// steps 36 / 96 (SACCH / FACCH1)
int nxdn_viterbi(int steps, const uint8_t* d_dibits, uint32_t n, uint32_t* d_words, uint32_t* d_metric, cudaStream_t st);
int pocsag_bch(uint32_t* d_words, uint8_t* d_ok, uint32_t n, cudaStream_t st);

}  // namespace test
}  // namespace dh
