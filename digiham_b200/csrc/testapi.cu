// testapi.cu — dh_test_*: device-level entry points for the FEC primitives (test hooks of the C ABI).
//
// the only known answers the reference ships are its own `corrections[]` tables; the parity tests
// therefore drive the DEVICE decoders exhaustively (all 2^n words for n <= 20, all <= t error patterns, random
// words, high-error trellis inputs that wrap the uint8 path metric) and compare them with the compiled reference
// word by word (tests/test_fec_gpu.py).  These calls take host buffers, run on the current device and are
// synchronous; the kernels call exactly the device functions the decoder kernels call.
#include "test_hooks.hpp"

namespace {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t bytes) {
        DH_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return DH_OK;
    }
};

}  // namespace

extern "C" {

int dh_test_fec(int code, uint32_t* h_words, uint8_t* h_ok, uint32_t n) {
    DH_REQUIRE(h_words != nullptr && h_ok != nullptr, DH_E_INVALID, "dh_test_fec: NULL buffer");
    DH_REQUIRE(code >= 0 && code <= 7, DH_E_INVALID, "dh_test_fec: unknown code %d", code);
    DevBuf w, ok;
    int rc = w.alloc((size_t) n * 4);
    if (rc == DH_OK) rc = ok.alloc(n);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(w.p, h_words, (size_t) n * 4, cudaMemcpyHostToDevice));
    if (code <= 5) rc = dh::test::dmr_fec(code, (uint32_t*) w.p, (uint8_t*) ok.p, n, nullptr);
    else if (code == 6) rc = dh::test::ysf_golay24((uint32_t*) w.p, (uint8_t*) ok.p, n, nullptr);
    else rc = dh::test::pocsag_bch((uint32_t*) w.p, (uint8_t*) ok.p, n, nullptr);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(h_words, w.p, (size_t) n * 4, cudaMemcpyDeviceToHost));
    DH_CUDA(cudaMemcpy(h_ok, ok.p, n, cudaMemcpyDeviceToHost));
    return DH_OK;
}

int dh_test_bptc(const uint8_t* h_payload, uint8_t* h_out, uint8_t* h_ok, uint32_t n) {
    DH_REQUIRE(h_payload != nullptr && h_out != nullptr && h_ok != nullptr, DH_E_INVALID, "dh_test_bptc: NULL buffer");
    DevBuf in, out, ok;
    int rc = in.alloc((size_t) n * 25);
    if (rc == DH_OK) rc = out.alloc((size_t) n * 12);
    if (rc == DH_OK) rc = ok.alloc(n);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(in.p, h_payload, (size_t) n * 25, cudaMemcpyHostToDevice));
    rc = dh::test::dmr_bptc((const uint8_t*) in.p, (uint8_t*) out.p, (uint8_t*) ok.p, n, nullptr);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(h_out, out.p, (size_t) n * 12, cudaMemcpyDeviceToHost));
    DH_CUDA(cudaMemcpy(h_ok, ok.p, n, cudaMemcpyDeviceToHost));
    return DH_OK;
}

int dh_test_viterbi(int variant, const uint8_t* h_dibits, uint32_t n, uint32_t* h_words, uint32_t* h_metric) {
    DH_REQUIRE(h_dibits != nullptr && h_words != nullptr && h_metric != nullptr, DH_E_INVALID,
               "dh_test_viterbi: NULL buffer");
    DH_REQUIRE(variant >= 0 && variant <= 3, DH_E_INVALID, "dh_test_viterbi: unknown variant %d", variant);
    const int steps = variant == 0 ? 100 : (variant == 1 ? 180 : (variant == 2 ? 36 : 96));
    const size_t nw = (size_t) (steps + 31) / 32;
    DevBuf in, words, metric;
    int rc = in.alloc((size_t) n * steps);
    if (rc == DH_OK) rc = words.alloc((size_t) n * nw * 4);
    if (rc == DH_OK) rc = metric.alloc((size_t) n * 4);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(in.p, h_dibits, (size_t) n * steps, cudaMemcpyHostToDevice));
    rc = variant <= 1 ? dh::test::ysf_viterbi(steps, (const uint8_t*) in.p, n, (uint32_t*) words.p, (uint32_t*) metric.p, nullptr)
                      : dh::test::nxdn_viterbi(steps, (const uint8_t*) in.p, n, (uint32_t*) words.p, (uint32_t*) metric.p, nullptr);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(h_words, words.p, (size_t) n * nw * 4, cudaMemcpyDeviceToHost));
    DH_CUDA(cudaMemcpy(h_metric, metric.p, (size_t) n * 4, cudaMemcpyDeviceToHost));
    return DH_OK;
}

}  // extern "C"
