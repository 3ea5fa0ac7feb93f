// common.cuh — shared host/device helpers of libdigiham_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string>

#include "../../include/digiham_b200.h"

namespace dh {

// ---- error plumbing -------------------------------------------------------------------------------------------
// Every C-ABI entry point returns DH_OK (0), a negative DH_E_* code, or a positive cudaError_t.
// The message of the last failure on the calling thread is kept for dh_last_error().
void set_error(const char* fmt, ...);

#define DH_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            dh::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int) e__;                                                                 \
        }                                                                                     \
    } while (0)

#define DH_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            dh::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        active = dev;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != active) cudaSetDevice(prev);
    }
    int active = -1;
};

int sm_count(int device);

// ---- per-channel state blobs (dh_*_state_export / _import) ---------------------------------------------------------
// Every blob starts with this header; import refuses a blob whose kind / channel count / configuration words differ
// from the bank it is loaded into.
struct StateHeader {
    uint32_t magic;      // 'DHST'
    uint32_t version;
    uint32_t kind;       // 1 RRC, 2 demodulator, 3 decoder, 4 pipe
    uint32_t channels;
    uint32_t cfg[4];     // bank configuration (taps / sps / protocol ...)
    uint64_t payload;    // bytes that follow the header
};
constexpr uint32_t kStateMagic = 0x54534844u;
constexpr uint32_t kStateVersion = 1;
inline StateHeader make_state_header(uint32_t kind, uint32_t channels, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                     uint64_t payload) {
    StateHeader h;
    h.magic = kStateMagic;
    h.version = kStateVersion;
    h.kind = kind;
    h.channels = channels;
    h.cfg[0] = c0;
    h.cfg[1] = c1;
    h.cfg[2] = c2;
    h.cfg[3] = c3;
    h.payload = payload;
    return h;
}
// DH_OK when `blob` is a state blob of exactly this configuration
int check_state_header(const void* blob, size_t bytes, const StateHeader& want, const char* who);

// ---- device-side PTX helpers (TMA 1-D bulk copies + mbarrier) -------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy (UBLKCP), completion signalled on an mbarrier. 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared -> global bulk copy, tracked by the bulk async-group of the issuing thread.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// wait until the source shared memory of all committed bulk stores has been read
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// order generic-proxy shared-memory writes before subsequent async-proxy (TMA) reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#endif  // __CUDACC__

}  // namespace dh
