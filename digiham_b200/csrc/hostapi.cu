// hostapi.cu — host-buffer variants of the per-stage entry points (dh_*_process_host).
//
// The bank kernels work on device blocks; these wrappers add the staging a caller with plain host memory needs
// (device scratch owned by a small per-handle cache, H2D copy, kernel, D2H copy, stream synchronisation).  They
// are what the header-compatible facade classes in include/*.hpp call, one module instance == a one-channel bank,
// so that an existing csdr pipe graph runs unchanged (BASELINE config 1, "plumbing").
#include "common.cuh"

#include <map>
#include <mutex>

namespace {

struct Scratch {
    void* a = nullptr;
    size_t a_bytes = 0;
    void* b = nullptr;
    size_t b_bytes = 0;
    void* c = nullptr;
    size_t c_bytes = 0;
};

std::mutex g_mutex;
std::map<const void*, Scratch> g_scratch;

int ensure(void** p, size_t* have, size_t need) {
    if (need <= *have) return DH_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    need = (need + need / 2 + 255) & ~(size_t) 255;
    DH_CUDA(cudaMalloc(p, need));
    *have = need;
    return DH_OK;
}

Scratch& scratch_of(const void* handle) {
    std::lock_guard<std::mutex> lock(g_mutex);
    return g_scratch[handle];
}

}  // namespace

extern "C" {

void dh_host_scratch_release(const void* handle) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_scratch.find(handle);
    if (it == g_scratch.end()) return;
    cudaFree(it->second.a);
    cudaFree(it->second.b);
    cudaFree(it->second.c);
    g_scratch.erase(it);
}

int dh_rrc_process_host(dh_rrc* h, uint32_t channels, const float* h_in, size_t in_pitch, float* h_out,
                        size_t out_pitch, size_t n) {
    DH_REQUIRE(h != nullptr && h_in != nullptr && h_out != nullptr, DH_E_INVALID, "dh_rrc_process_host: NULL argument");
    DH_REQUIRE(channels == dh_rrc_channels(h), DH_E_INVALID, "dh_rrc_process_host: channels=%u but the bank has %u",
               channels, dh_rrc_channels(h));
    if (n == 0) return DH_OK;
    DH_REQUIRE(in_pitch >= n && out_pitch >= n, DH_E_INVALID, "dh_rrc_process_host: pitch < n");
    Scratch& s = scratch_of(h);
    const size_t pitch = (n + 3) & ~(size_t) 3;
    int rc = ensure(&s.a, &s.a_bytes, (size_t) channels * pitch * sizeof(float));
    if (rc == DH_OK) rc = ensure(&s.b, &s.b_bytes, (size_t) channels * pitch * sizeof(float));
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy2D(s.a, pitch * sizeof(float), h_in, in_pitch * sizeof(float), n * sizeof(float), channels,
                         cudaMemcpyHostToDevice));
    rc = dh_rrc_process(h, (const float*) s.a, pitch, (float*) s.b, pitch, n, nullptr);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy2D(h_out, out_pitch * sizeof(float), s.b, pitch * sizeof(float), n * sizeof(float), channels,
                         cudaMemcpyDeviceToHost));
    return DH_OK;
}

int dh_demod_process_host(dh_demod* h, uint32_t channels, const float* h_in, size_t in_pitch, size_t n,
                          uint8_t* h_sym, size_t sym_pitch, uint32_t* h_nsym) {
    DH_REQUIRE(h != nullptr && h_sym != nullptr && h_nsym != nullptr, DH_E_INVALID,
               "dh_demod_process_host: NULL argument");
    DH_REQUIRE(channels == dh_demod_channels(h), DH_E_INVALID, "dh_demod_process_host: channels=%u but the bank has %u",
               channels, dh_demod_channels(h));
    if (n == 0) {
        for (uint32_t c = 0; c < channels; c++) h_nsym[c] = 0;
        return DH_OK;
    }
    DH_REQUIRE(h_in != nullptr && in_pitch >= n, DH_E_INVALID, "dh_demod_process_host: bad input");
    const size_t need_syms = dh_demod_max_symbols(h, n);
    DH_REQUIRE(sym_pitch >= need_syms, DH_E_INVALID,
               "dh_demod_process_host: sym_pitch %zu too small, need dh_demod_max_symbols(n) = %zu", sym_pitch, need_syms);
    float* d_in = nullptr;
    size_t d_pitch = 0;
    int rc = dh_demod_reserve(h, n, &d_in, &d_pitch);
    if (rc != DH_OK) return rc;
    Scratch& s = scratch_of(h);
    rc = ensure(&s.a, &s.a_bytes, (size_t) channels * need_syms);
    if (rc == DH_OK) rc = ensure(&s.b, &s.b_bytes, (size_t) channels * sizeof(uint32_t));
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy2D(d_in, d_pitch * sizeof(float), h_in, in_pitch * sizeof(float), n * sizeof(float), channels,
                         cudaMemcpyHostToDevice));
    rc = dh_demod_process(h, d_in, d_pitch, n, (uint8_t*) s.a, need_syms, (uint32_t*) s.b, nullptr);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy(h_nsym, s.b, (size_t) channels * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    uint32_t mx = 0;
    for (uint32_t c = 0; c < channels; c++) mx = h_nsym[c] > mx ? h_nsym[c] : mx;
    if (mx) DH_CUDA(cudaMemcpy2D(h_sym, sym_pitch, s.a, need_syms, mx, channels, cudaMemcpyDeviceToHost));
    return DH_OK;
}

int dh_decoder_process_host(dh_decoder* h, uint32_t channels, const uint8_t* h_sym, size_t sym_pitch,
                            const uint32_t* h_nsym) {
    DH_REQUIRE(h != nullptr && h_nsym != nullptr, DH_E_INVALID, "dh_decoder_process_host: NULL argument");
    DH_REQUIRE(channels == dh_decoder_channels(h), DH_E_INVALID,
               "dh_decoder_process_host: channels=%u but the bank has %u", channels, dh_decoder_channels(h));
    uint32_t mx = 0;
    for (uint32_t c = 0; c < channels; c++) mx = h_nsym[c] > mx ? h_nsym[c] : mx;
    if (mx == 0) return DH_OK;
    DH_REQUIRE(h_sym != nullptr && sym_pitch >= mx, DH_E_INVALID, "dh_decoder_process_host: bad symbol buffer");
    uint8_t* d_sym = nullptr;
    size_t d_pitch = 0;
    int rc = dh_decoder_reserve(h, mx, &d_sym, &d_pitch);
    if (rc != DH_OK) return rc;
    Scratch& s = scratch_of(h);
    rc = ensure(&s.b, &s.b_bytes, (size_t) channels * sizeof(uint32_t));
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy2D(d_sym, d_pitch, h_sym, sym_pitch, mx, channels, cudaMemcpyHostToDevice));
    DH_CUDA(cudaMemcpy(s.b, h_nsym, (size_t) channels * sizeof(uint32_t), cudaMemcpyHostToDevice));
    rc = dh_decoder_process(h, d_sym, d_pitch, (const uint32_t*) s.b, mx, nullptr);
    if (rc != DH_OK) return rc;
    return dh_decoder_collect(h, nullptr);
}

int dh_dvf_process_host(dh_dvf* h, uint32_t channels, const int16_t* h_in, size_t in_pitch, int16_t* h_out,
                        size_t out_pitch, size_t n) {
    DH_REQUIRE(h != nullptr && h_in != nullptr && h_out != nullptr, DH_E_INVALID, "dh_dvf_process_host: NULL argument");
    DH_REQUIRE(channels == dh_dvf_channels(h), DH_E_INVALID, "dh_dvf_process_host: channels=%u but the bank has %u",
               channels, dh_dvf_channels(h));
    if (n == 0) return DH_OK;
    DH_REQUIRE(in_pitch >= n && out_pitch >= n, DH_E_INVALID, "dh_dvf_process_host: pitch < n");
    Scratch& s = scratch_of(h);
    int rc = ensure(&s.a, &s.a_bytes, (size_t) channels * n * sizeof(int16_t));
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy2D(s.a, n * sizeof(int16_t), h_in, in_pitch * sizeof(int16_t), n * sizeof(int16_t), channels,
                         cudaMemcpyHostToDevice));
    rc = dh_dvf_process(h, (const int16_t*) s.a, n, (int16_t*) s.a, n, n, nullptr);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemcpy2D(h_out, out_pitch * sizeof(int16_t), s.a, n * sizeof(int16_t), n * sizeof(int16_t), channels,
                         cudaMemcpyDeviceToHost));
    return DH_OK;
}

}  // extern "C"
