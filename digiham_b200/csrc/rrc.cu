// rrc.cu — K1: root-raised-cosine FIR bank (dh_rrc_*), sm_100a.
//
// Replaces Digiham::RrcFilter::RrcFilter::process/filter (reference src/rrc_filter/rrc_filter.cpp:16-34) for N
// channels at once.  Bit-exact contract: per output sample
//     sum = 0.0f; for i in 0..nZeros: sum = fl32(sum + fl32(c[i] * x[t - nZeros + i]));  out = fl32(fl64(sum) / gain)
// i.e. separately rounded products and a strictly ordered float32 accumulation (the x86-64 reference build has
// no FMA), then one double division rounded to float.  For the two built-in gains the division is replaced by a
// multiplication with the correctly rounded reciprocal: tests/test_host_logic.py (test_reciprocal_gain_exhaustive) proves, by exhaustion over all
// 2^32 float inputs, that fl32(fl64(s) * fl64(1/g)) == fl32(fl64(s) / g) for g in {8.337797030, 16.67711971}.
//
// Kernel shape (1-D convolution, FP32-issue bound at 2*(nZeros+1) flop per sample — no tensor cores):
//   * one CTA = one (channel, time tile) of TILE = 128 threads x R outputs; R is odd (13..25, chosen per call so
//     that the stream splits into equal tiles) so that the per-thread windows (stride R floats) fall into 32
//     different shared-memory banks without any padding;
//   * the tile and its nZeros-sample halo are brought in by TMA 1-D bulk copies (cp.async.bulk -> UBLKCP) that
//     signal an mbarrier; outputs leave through shared memory and one bulk store, so every HBM access is a
//     full-line burst and the SM issue slots are left to FMUL/FADD;
//   * each thread keeps R accumulators and an R-deep sliding window in registers: per tap it issues R FMUL +
//     R FADD + one LDS (the window advances by one sample) + one constant-bank tap fetch;
//   * latency is hidden by occupancy (≈9 KB smem, <64 regs per thread -> many CTAs per SM), not by an
//     intra-CTA pipeline.
#include "common.cuh"
#include "tables.inc"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

constexpr int kThreads = 128;
constexpr int kRDefault = 17;         // outputs per thread; odd, see the kernel comment
constexpr int kMaxR = 25;
constexpr int kMaxTile = kThreads * kMaxR;
constexpr int kMaxZeros = 1024;
constexpr int kMaxTapsParam = 164;    // taps that travel in the kernel parameter (constant bank 0)
// The parameter copy is laid out in groups of R taps padded to a multiple of 4 floats, so that the rolled loop
// fetches the taps of one group with 16-byte uniform loads (LDCU.128) instead of one LDCU per tap.
constexpr int kTapSlots = 224;        // >= groups * pad4(R) for 164 taps at every R in 13..25

struct alignas(16) TapBlock {
    float c[kTapSlots];
};
__host__ __device__ constexpr int pad4(int r) { return (r + 3) & ~3; }

struct RrcParams {
    const void* in;         // float32 rows, or int16 rows (S16 kernels: dh_rrc_process_s16)
    float* out;
    const float* hist_in;   // [channels][nz]  last nz inputs before this call
    float* hist_out;        // [channels][nz]  last nz inputs after this call
    const float* taps_g;    // taps in global memory (only used when nz + 1 > kMaxTapsParam)
    unsigned long long in_pitch;
    unsigned long long out_pitch;
    double gain;            // divisor, or its reciprocal when mul_recip != 0
    int n;
    int nz;
    int tiles;
    int pad_;
};

// Outputs per thread for a call of n samples.  A CTA slot costs the same whether its tile is full or ragged (the
// surviving warps of a ragged tile run no faster), so the tile size 128 * R is chosen from odd R in 13..25 to
// minimise tiles * (work per tile): e.g. n = 48000 -> R = 25, 15 equal tiles of 3200 samples (fewest window loads,
// tap fetches and prologues per output) instead of 22 + a ragged one at R = 17; R = 15, 25 equal tiles, when the
// smaller register footprint is preferred (dh_rrc_set_tile_preference).
inline int pick_r(size_t n, int nz, bool prefer_small) {
    static const int forced = [] {   // tuning switch, read once
        const char* env = getenv("DH_RRC_R");
        const int r = env ? atoi(env) : 0;
        return (r >= 13 && r <= kMaxR && (r & 1)) ? r : 0;
    }();
    if (forced) return forced;
    int best = kRDefault;
    double best_cost = 1e300;
    // beside other kernels the large tiles (R > 19: 53..63 registers) lose more through their footprint than they gain
    for (int r = prefer_small ? 19 : kMaxR; r >= 13; r -= 2) {
        const size_t tiles = (n + (size_t) kThreads * r - 1) / ((size_t) kThreads * r);
        // per tile and thread: (nz + 1) * (2 r + 2) issue slots + fixed prologue / epilogue
        double cost = (double) tiles * ((nz + 1) * (2.0 * r + 2.0) + 6.0 * r + 150.0);
        // beside other kernels (dh_pipe_set_async) a smaller register footprint keeps more CTAs resident: measured
        // 1.334 ms per pipelined DMR step at R = 15 (40 registers) against 1.352 at R = 19 and 1.370 at R = 25 (63
        // registers), although R = 25 is 3 % faster alone (0.979 vs 1.008 ms, profiles/r02_sweep_step.txt)
        if (prefer_small) cost *= 1.0 + 0.004 * r;
        if (cost < best_cost * 0.999) {
            best_cost = cost;
            best = r;
        }
    }
    return best;
}

// `csdr convert -i s16 -o float` (the step in front of rrc_filter in reference examples/dmr-decoder.sh:13-15):
// out = (float) in / SHRT_MAX, one IEEE float32 division.  Computed as q = f * r, q' = fma(fma(-q, 32767, f), r, q)
// with r = fl32(1 / 32767): proven equal to the division for all 65536 inputs (tests/test_host_logic.py).
__device__ __forceinline__ float s16_to_float(int v) {
    const float f = (float) v;
    const float r = 1.0f / 32767.0f;
    const float q = __fmul_rn(f, r);
    return __fmaf_rn(__fmaf_rn(-q, 32767.0f, f), r, q);
}

// NZ_CT > 0: compile-time tap count; RECIP: scale by the reciprocal gain (built-in filters) instead of dividing;
// S16: the input rows are int16 and are converted while the tile sits in shared memory (fused csdr convert)
template <int NZ_CT, bool RECIP, int kR, bool S16>
__global__ void __launch_bounds__(kThreads) rrc_fir_kernel(const __grid_constant__ RrcParams p,
                                                          const __grid_constant__ TapBlock taps) {
    constexpr int kTile = kThreads * kR;
    extern __shared__ __align__(128) float s[];   // [nz + kTile] inputs, later reused for kTile outputs
    __shared__ __align__(8) uint64_t bar;

    const int nz = NZ_CT > 0 ? NZ_CT : p.nz;
    const int tile = blockIdx.x;   // 2-D grid (tiles, channels): no integer division in the prologue
    const int ch = blockIdx.y;
    const int t0 = tile * kTile;
    const int valid = min(kTile, p.n - t0);
    const float* in_row = static_cast<const float*>(p.in) + (size_t) ch * p.in_pitch;
    const int16_t* in_row16 = static_cast<const int16_t*>(p.in) + (size_t) ch * p.in_pitch;
    float* out_row = p.out + (size_t) ch * p.out_pitch;
    const int tid = threadIdx.x;
    // S16: the raw int16 tile lands in the upper half of the float buffer (bytes [2F, 4F), F = nz + kTile) and is
    // converted into place through registers; the FIR history stays float32 (tile 0 gets its halo as floats)
    const int F = nz + kTile;
    int16_t* raw = reinterpret_cast<int16_t*>(s) + F;

    if (tid == 0) {
        dh::mbar_init(&bar, 1);
        dh::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        if (S16) {
            const uint32_t main_bytes = (uint32_t) ((valid + 7) & ~7) * 2u;
            const uint32_t halo_bytes = tile == 0 ? (uint32_t) nz * 4u : (uint32_t) nz * 2u;
            dh::mbar_expect_tx(&bar, main_bytes + halo_bytes);
            if (tile == 0) dh::bulk_g2s(s, p.hist_in + (size_t) ch * nz, halo_bytes, &bar);
            else dh::bulk_g2s(raw, in_row16 + t0 - nz, halo_bytes, &bar);
            dh::bulk_g2s(raw + nz, in_row16 + t0, main_bytes, &bar);
        } else {
            const uint32_t main_bytes = (uint32_t) ((valid + 3) & ~3) * 4u;
            const uint32_t halo_bytes = (uint32_t) nz * 4u;
            dh::mbar_expect_tx(&bar, main_bytes + halo_bytes);
            const float* halo = tile == 0 ? p.hist_in + (size_t) ch * nz : in_row + t0 - nz;
            dh::bulk_g2s(s, halo, halo_bytes, &bar);
            dh::bulk_g2s(s + nz, in_row + t0, main_bytes, &bar);
        }
    }

    // carry the FIR history across calls: the CTA of the last tile publishes the last nz inputs of the stream
    if (tile == p.tiles - 1) {
        const float* hin = p.hist_in + (size_t) ch * nz;
        float* hout = p.hist_out + (size_t) ch * nz;
        for (int j = tid; j < nz; j += kThreads) {
            int idx = p.n - nz + j;
            hout[j] = idx >= 0 ? (S16 ? s16_to_float(in_row16[idx]) : in_row[idx]) : hin[nz + idx];
        }
    }

    // taps that do not fit the parameter block are staged behind the sample tile
    const bool taps_in_smem = NZ_CT == 0 && nz + 1 > kMaxTapsParam;
    float* s_taps = s + nz + kTile;
    if (taps_in_smem) {
        for (int j = tid; j <= nz; j += kThreads) s_taps[j] = p.taps_g[j];
        __syncthreads();
    }

    dh::mbar_wait(&bar, 0);

    if (S16) {
        // pairs of samples: one 32-bit shared load, two conversions, one 64-bit store; a float pair only overwrites
        // raw samples of lower or equal index, and every thread holds its raw words in registers before anybody writes
        constexpr int kMaxPairs = (kMaxZeros + kTile) / 2;
        constexpr int kIter = (kMaxPairs + kThreads - 1) / kThreads;
        const uint32_t* raw32 = reinterpret_cast<const uint32_t*>(raw);
        const int first = tile == 0 ? nz / 2 : 0;                      // tile 0: the halo already is float32
        const int last = (nz + ((valid + 7) & ~7)) / 2;                 // pairs [first, last)
        constexpr int kIterCt = NZ_CT > 0 ? ((NZ_CT + kTile) / 2 + kThreads - 1) / kThreads : kIter;
        uint32_t held[kIterCt];
#pragma unroll
        for (int m = 0; m < kIterCt; m++) {
            const int i = first + tid + m * kThreads;
            held[m] = i < last ? raw32[i] : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < kIterCt; m++) {
            const int i = first + tid + m * kThreads;
            if (i < last) {
                float2 v;
                v.x = s16_to_float((int) (short) (held[m] & 0xffffu));
                v.y = s16_to_float((int) (short) (held[m] >> 16));
                reinterpret_cast<float2*>(s)[i] = v;
            }
        }
        __syncthreads();
    }

    // s[j] = x[t0 - nz + j]; output r of this thread is sample t0 + base + r and needs s[base + r + i], i = 0..nz
    const int base = tid * kR;
    const float* sw = s + base;
    float acc[kR], w[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) {
        acc[r] = 0.0f;
        w[r] = sw[r];
    }

    // a ragged last tile: warps whose outputs all lie beyond the end of the stream skip the arithmetic
    // (warp-uniform, so no divergence inside the FMUL/FADD stream)
    const bool warp_live = (tid & ~31) * kR < valid;
    if (warp_live) {
        const int ntaps = nz + 1;
        const int full = ntaps / kR;   // compile-time for the built-in filters
        int i0 = 0;
        // rolled loop over groups of kR taps: ~10 KB of code.  Fully unrolling the 81-tap filter (45 KB, taps fetched
        // by LDCU.128) was measured 9 % slower on B200 (instruction-cache misses).
#pragma unroll 1
        for (int it = 0; it < full; it++, i0 += kR) {
            const float* nxt = sw + i0 + kR;
#pragma unroll
            for (int k = 0; k < kR; k++) {
                const float c = taps_in_smem ? s_taps[i0 + k] : taps.c[it * pad4(kR) + k];
#pragma unroll
                for (int r = 0; r < kR; r++) acc[r] = __fadd_rn(acc[r], __fmul_rn(c, w[(r + k) % kR]));
                // slot k held s[base + i0 + k] (the oldest sample, last used by r = 0); it now receives the
                // sample that output r = kR-1 needs at the next tap.  Never read past tap index nz.
                if (i0 + k < nz) w[k] = nxt[k];
            }
        }
        const int rem = ntaps - i0;   // < kR, identical for all threads
        const float* nxt = sw + i0 + kR;
#pragma unroll
        for (int k = 0; k < kR - 1; k++) {
            if (k < rem) {
                const float c = taps_in_smem ? s_taps[i0 + k] : taps.c[full * pad4(kR) + k];
#pragma unroll
                for (int r = 0; r < kR; r++) acc[r] = __fadd_rn(acc[r], __fmul_rn(c, w[(r + k) % kR]));
                if (i0 + k < nz) w[k] = nxt[k];
            }
        }
    }

    // everybody is done reading the input tile: reuse it for the outputs
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kR; r++) {
        const double q = RECIP ? __dmul_rn((double) acc[r], p.gain) : __ddiv_rn((double) acc[r], p.gain);
        s[base + r] = __double2float_rn(q);
    }
    dh::fence_async_smem();
    __syncthreads();

    const int bulk_elems = valid & ~3;
    if (tid == 0 && bulk_elems > 0) {
        dh::bulk_s2g(out_row + t0, s, (uint32_t) bulk_elems * 4u);
        dh::bulk_commit();
        dh::bulk_wait_read0();
    }
    // at most three trailing samples of the stream do not fill a 16-byte unit
    if (tid < valid - bulk_elems) out_row[t0 + bulk_elems + tid] = s[bulk_elems + tid];
}

}  // namespace

struct dh_rrc {
    int device = 0;
    uint32_t channels = 0;
    int nz = 0;
    double gain = 1.0;
    int mul_recip = 0;
    float flat_taps[kMaxTapsParam] = {0};   // taps 0..min(nz, kMaxTapsParam - 1) in order
    TapBlock taps{};                        // grouped layout for `taps_r` outputs per thread
    int taps_r = 0;
    float* d_taps = nullptr;   // global copy, used for long custom filters
    float* d_hist = nullptr;   // [2][channels][nz]
    int cur = 0;
    int prefer_small_tiles = 0;
};

namespace {

int rrc_build(dh_rrc** out, int device, uint32_t channels, uint32_t nz, double gain, int mul_recip,
              const float* coeffs) {
    DH_REQUIRE(out != nullptr, DH_E_INVALID, "dh_rrc_create: out is NULL");
    *out = nullptr;
    DH_REQUIRE(channels > 0, DH_E_INVALID, "dh_rrc_create: channels must be > 0");
    DH_REQUIRE(nz >= 4 && nz % 4 == 0 && nz <= kMaxZeros, DH_E_UNSUPPORTED,
               "dh_rrc_create: nZeros=%u not supported (multiple of 4, 4..%d)", nz, kMaxZeros);
    DH_REQUIRE(coeffs != nullptr, DH_E_INVALID, "dh_rrc_create: coeffs is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        dh::set_error("dh_rrc_create: no CUDA device available (this library has no CPU fallback)");
        return DH_E_NODEVICE;
    }
    DH_REQUIRE(device >= 0 && device < ndev, DH_E_INVALID, "dh_rrc_create: device %d out of range", device);
    dh::DeviceGuard guard(device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", device);
    dh_rrc* h = new (std::nothrow) dh_rrc();
    DH_REQUIRE(h != nullptr, DH_E_NOMEM, "dh_rrc_create: out of host memory");
    h->device = device;
    h->channels = channels;
    h->nz = (int) nz;
    h->gain = mul_recip ? 1.0 / gain : gain;
    h->mul_recip = mul_recip;
    for (uint32_t i = 0; i <= nz && i < (uint32_t) kMaxTapsParam; i++) h->flat_taps[i] = coeffs[i];
    size_t hist_bytes = 2 * (size_t) channels * nz * sizeof(float);
    cudaError_t e = cudaMalloc(&h->d_hist, hist_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->d_hist, 0, hist_bytes);
    if (e == cudaSuccess && nz + 1 > (uint32_t) kMaxTapsParam) {
        e = cudaMalloc(&h->d_taps, (nz + 1) * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_taps, coeffs, (nz + 1) * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        dh::set_error("dh_rrc_create: %s", cudaGetErrorString(e));
        cudaFree(h->d_hist);
        cudaFree(h->d_taps);
        delete h;
        return (int) e;
    }
    *out = h;
    return DH_OK;
}

void bits_to_floats(const uint32_t* bits, int n, std::vector<float>& out) {
    out.resize(n);
    std::memcpy(out.data(), bits, n * sizeof(float));
}

}  // namespace

extern "C" {

int dh_rrc_create(dh_rrc** out, int device, uint32_t channels, int kind) {
    std::vector<float> taps;
    if (kind == DH_RRC_WIDE) {
        bits_to_floats(dh_rrc_wide_taps_bits, DH_RRC_WIDE_NZEROS + 1, taps);
        return rrc_build(out, device, channels, DH_RRC_WIDE_NZEROS, DH_RRC_WIDE_GAIN, 1, taps.data());
    }
    if (kind == DH_RRC_NARROW) {
        bits_to_floats(dh_rrc_narrow_taps_bits, DH_RRC_NARROW_NZEROS + 1, taps);
        return rrc_build(out, device, channels, DH_RRC_NARROW_NZEROS, DH_RRC_NARROW_GAIN, 1, taps.data());
    }
    if (out) *out = nullptr;
    dh::set_error("dh_rrc_create: unknown kind %d", kind);
    return DH_E_INVALID;
}

int dh_rrc_create_custom(dh_rrc** out, int device, uint32_t channels, uint32_t n_zeros, double gain,
                         const float* h_coeffs) {
    return rrc_build(out, device, channels, n_zeros, gain, 0, h_coeffs);
}

}  // extern "C"

namespace {

int rrc_launch(dh_rrc* h, const void* d_in, size_t in_pitch, float* d_out, size_t out_pitch, size_t n, void* stream,
               bool s16, const char* who) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "%s: handle is NULL", who);
    if (n == 0) return DH_OK;
    DH_REQUIRE(d_in != nullptr && d_out != nullptr, DH_E_INVALID, "%s: NULL buffer", who);
    DH_REQUIRE(d_in != (const void*) d_out, DH_E_INVALID, "%s: in-place operation is not supported", who);
    DH_REQUIRE(((uintptr_t) d_in % 16 == 0) && ((uintptr_t) d_out % 16 == 0), DH_E_INVALID,
               "%s: buffers must be 16-byte aligned", who);
    const size_t n4 = (n + 3) & ~(size_t) 3;
    const size_t in_unit = s16 ? 8 : 4;   // 16-byte rows
    const size_t n_in = (n + in_unit - 1) / in_unit * in_unit;
    DH_REQUIRE(in_pitch % in_unit == 0 && out_pitch % 4 == 0 && in_pitch >= n_in && out_pitch >= n4, DH_E_INVALID,
               "%s: pitches must be multiples of %zu (in) / 4 (out) and >= n rounded up (n=%zu in=%zu out=%zu)", who,
               in_unit, n, in_pitch, out_pitch);
    DH_REQUIRE(!s16 || h->nz % 8 == 0, DH_E_UNSUPPORTED, "%s: int16 input needs nZeros %% 8 == 0", who);
    DH_REQUIRE(n <= 0x7fffffffu - kMaxTile, DH_E_INVALID, "%s: n too large", who);
    const int r = pick_r(n, h->nz, h->prefer_small_tiles != 0);
    const size_t tile = (size_t) kThreads * r;
    const size_t tiles = (n + tile - 1) / tile;
    DH_REQUIRE(tiles <= 0x7fffffffu, DH_E_INVALID, "%s: too many tiles", who);

    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "%s: cannot switch to device %d", who, h->device);
    RrcParams p;
    p.in = d_in;
    p.out = d_out;
    const size_t hist_elems = (size_t) h->channels * h->nz;
    p.hist_in = h->d_hist + (size_t) h->cur * hist_elems;
    p.hist_out = h->d_hist + (size_t) (h->cur ^ 1) * hist_elems;
    p.taps_g = h->d_taps;
    p.in_pitch = in_pitch;
    p.out_pitch = out_pitch;
    p.gain = h->gain;
    p.n = (int) n;
    p.nz = h->nz;
    p.tiles = (int) tiles;
    p.pad_ = 0;

    cudaStream_t st = (cudaStream_t) stream;
    if (h->taps_r != r) {   // (re)build the grouped parameter layout for this R
        std::memset(&h->taps, 0, sizeof(h->taps));
        for (int i = 0; i <= h->nz && i < kMaxTapsParam; i++) h->taps.c[(i / r) * pad4(r) + i % r] = h->flat_taps[i];
        h->taps_r = r;
    }
    size_t smem = (size_t) (h->nz + tile) * sizeof(float);
    // experiment switch: extra (unused) dynamic shared memory per CTA caps the resident CTAs per SM, to measure how much
    // of the slowdown beside K2 / K3 is lost occupancy
    static const size_t extra_smem = getenv("DH_RRC_EXTRA_SMEM") ? (size_t) atoi(getenv("DH_RRC_EXTRA_SMEM")) : 0;
    if (smem + extra_smem <= 48 * 1024) smem += extra_smem;
    const int which = h->nz == 80 && h->mul_recip ? 0 : (h->nz == 160 && h->mul_recip ? 1 : 2);
    if (which == 2 && h->nz + 1 > kMaxTapsParam) smem += (size_t) (h->nz + 1) * sizeof(float);
#define DH_LAUNCH_RRC2(RR, SS)                                                                          \
    do {                                                                                                \
        if (which == 0) rrc_fir_kernel<80, true, RR, SS><<<grid, kThreads, smem, st>>>(q, h->taps);      \
        else if (which == 1) rrc_fir_kernel<160, true, RR, SS><<<grid, kThreads, smem, st>>>(q, h->taps); \
        else rrc_fir_kernel<0, false, RR, SS><<<grid, kThreads, smem, st>>>(q, h->taps);                 \
    } while (0)
#define DH_LAUNCH_RRC(RR)                    \
    do {                                     \
        if (s16) DH_LAUNCH_RRC2(RR, true);   \
        else DH_LAUNCH_RRC2(RR, false);      \
    } while (0)
    // grid = (tiles, channels); grid.y is limited to 65535, larger banks are launched in channel slices
    const size_t in_elem = s16 ? sizeof(int16_t) : sizeof(float);
    for (size_t c0 = 0; c0 < h->channels; c0 += 65535) {
        const size_t cnt = std::min<size_t>(65535, h->channels - c0);
        RrcParams q = p;
        q.in = static_cast<const char*>(p.in) + c0 * in_pitch * in_elem;
        q.out += c0 * out_pitch;
        q.hist_in += c0 * h->nz;
        q.hist_out += c0 * h->nz;
        const dim3 grid((unsigned) tiles, (unsigned) cnt, 1);
        switch (r) {
            case 13: DH_LAUNCH_RRC(13); break;
            case 15: DH_LAUNCH_RRC(15); break;
            case 19: DH_LAUNCH_RRC(19); break;
            case 21: DH_LAUNCH_RRC(21); break;
            case 23: DH_LAUNCH_RRC(23); break;
            case 25: DH_LAUNCH_RRC(25); break;
            default: DH_LAUNCH_RRC(17); break;
        }
    }
#undef DH_LAUNCH_RRC
#undef DH_LAUNCH_RRC2
    DH_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return DH_OK;
}

}  // namespace

extern "C" {

int dh_rrc_process(dh_rrc* h, const float* d_in, size_t in_pitch, float* d_out, size_t out_pitch, size_t n,
                   void* stream) {
    return rrc_launch(h, d_in, in_pitch, d_out, out_pitch, n, stream, false, "dh_rrc_process");
}

int dh_rrc_process_s16(dh_rrc* h, const int16_t* d_in, size_t in_pitch, float* d_out, size_t out_pitch, size_t n,
                       void* stream) {
    return rrc_launch(h, d_in, in_pitch, d_out, out_pitch, n, stream, true, "dh_rrc_process_s16");
}

uint32_t dh_rrc_channels(const dh_rrc* h) { return h ? h->channels : 0; }

int dh_rrc_set_tile_preference(dh_rrc* h, int prefer_small) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_rrc_set_tile_preference: handle is NULL");
    h->prefer_small_tiles = prefer_small != 0;
    return DH_OK;
}

// ---- state: the FIR history (the last nZeros inputs of every channel) ----------------------------------------------
namespace {
dh::StateHeader rrc_header(const dh_rrc* h) {
    uint32_t g[2];
    std::memcpy(g, &h->gain, sizeof(g));
    return dh::make_state_header(1, h->channels, (uint32_t) h->nz, (uint32_t) h->mul_recip, g[0], g[1],
                                 (uint64_t) h->channels * h->nz * sizeof(float));
}
}  // namespace

int dh_rrc_state_size(const dh_rrc* h, size_t* bytes) {
    DH_REQUIRE(h != nullptr && bytes != nullptr, DH_E_INVALID, "dh_rrc_state_size: NULL argument");
    *bytes = sizeof(dh::StateHeader) + (size_t) h->channels * h->nz * sizeof(float);
    return DH_OK;
}

int dh_rrc_state_export(dh_rrc* h, void* h_buf, size_t cap, size_t* written, void* stream) {
    DH_REQUIRE(h != nullptr && h_buf != nullptr, DH_E_INVALID, "dh_rrc_state_export: NULL argument");
    const dh::StateHeader hd = rrc_header(h);
    DH_REQUIRE(cap >= sizeof(hd) + hd.payload, DH_E_INVALID, "dh_rrc_state_export: buffer too small");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    std::memcpy(h_buf, &hd, sizeof(hd));
    const float* cur = h->d_hist + (size_t) h->cur * h->channels * h->nz;
    DH_CUDA(cudaMemcpyAsync(static_cast<char*>(h_buf) + sizeof(hd), cur, hd.payload, cudaMemcpyDeviceToHost,
                            (cudaStream_t) stream));
    DH_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    if (written) *written = sizeof(hd) + hd.payload;
    return DH_OK;
}

int dh_rrc_state_import(dh_rrc* h, const void* h_buf, size_t bytes, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_rrc_state_import: handle is NULL");
    const dh::StateHeader hd = rrc_header(h);
    int rc = dh::check_state_header(h_buf, bytes, hd, "dh_rrc_state_import");
    if (rc != DH_OK) return rc;
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    float* cur = h->d_hist + (size_t) h->cur * h->channels * h->nz;
    DH_CUDA(cudaMemcpyAsync(cur, static_cast<const char*>(h_buf) + sizeof(hd), hd.payload, cudaMemcpyHostToDevice,
                            (cudaStream_t) stream));
    DH_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    return DH_OK;
}

int dh_rrc_reset(dh_rrc* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_rrc_reset: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    DH_CUDA(cudaMemsetAsync(h->d_hist, 0, 2 * (size_t) h->channels * h->nz * sizeof(float), (cudaStream_t) stream));
    h->cur = 0;
    return DH_OK;
}

void dh_rrc_destroy(dh_rrc* h) {
    if (!h) return;
    dh::DeviceGuard guard(h->device);
    cudaFree(h->d_hist);
    cudaFree(h->d_taps);
    delete h;
}

}  // extern "C"
