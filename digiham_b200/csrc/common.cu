// common.cu — error plumbing and device queries shared by all banks of libdigiham_b200.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace dh {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        return 148;
    }
    return n;
}

int check_state_header(const void* blob, size_t bytes, const StateHeader& want, const char* who) {
    DH_REQUIRE(blob != nullptr && bytes >= sizeof(StateHeader), DH_E_INVALID, "%s: blob too small for a state header", who);
    StateHeader got;
    memcpy(&got, blob, sizeof(got));
    DH_REQUIRE(got.magic == kStateMagic && got.version == kStateVersion, DH_E_INVALID,
               "%s: not a state blob of this library version", who);
    DH_REQUIRE(got.kind == want.kind && got.channels == want.channels && memcmp(got.cfg, want.cfg, sizeof(got.cfg)) == 0,
               DH_E_INVALID, "%s: the blob belongs to a different bank (kind %u/%u, channels %u/%u, configuration)", who,
               got.kind, want.kind, got.channels, want.channels);
    DH_REQUIRE(bytes >= sizeof(StateHeader) + got.payload, DH_E_INVALID, "%s: truncated blob", who);
    return DH_OK;
}

}  // namespace dh

extern "C" {

const char* dh_last_error(void) { return dh::g_error; }

const char* dh_version(void) { return "0.1.0-b200"; }

int dh_host_alloc(void** ptr, size_t bytes, int write_combined) {
    DH_REQUIRE(ptr != nullptr && bytes > 0, DH_E_INVALID, "dh_host_alloc: bad argument");
    *ptr = nullptr;
    DH_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return DH_OK;
}

int dh_host_free(void* ptr) {
    if (!ptr) return DH_OK;
    DH_CUDA(cudaFreeHost(ptr));
    return DH_OK;
}

int dh_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        if (count) *count = 0;
        dh::set_error("no CUDA device available: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return DH_E_NODEVICE;
    }
    if (count) *count = n;
    return DH_OK;
}

}
