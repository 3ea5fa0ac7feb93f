// dvf.cu — K6: DigitalVoiceFilter bank (dh_dvf_*), sm_100a.
//
// Replaces Digiham::DigitalVoice::DigitalVoiceFilter::process/filter (reference
// src/digitalvoice_filter/digitalvoice_filter.cpp:6-45, include/digitalvoice_filter.hpp:12-19): a 10th-order
// Butterworth band-pass IIR on 8 kHz int16 audio, run after the AMBE vocoder — NOT part of the RRC->GFSK->DMR
// pipe; it is a stand-alone module.
//
// Bit-exact contract per sample (x86-64 reference build, no FMA):
//   x      = fl32(fl32(in) / 32767.0f);  xv[10] = fl32(x / 5.0f)                      (GAIN hand-tuned to 5)
//   f      = fl32(fl32(fl32(xv10 - xv0) + fl32(5 * fl32(xv2 - xv8))) + fl32(10 * fl32(xv6 - xv4)))   [float]
//   acc    = fl64(f); for j = 0..9: acc = fl64(acc + fl64(c_j * fl64(yv[j])))          [double, left to right]
//   yv[10] = fl32(acc);  out = (short) fl32(yv[10] * 32767.0f)   (cvttss2si, low 16 bits)
// The order-10 feedback makes time strictly sequential per channel: one thread owns one channel; 32 channels x 64
// samples tiles go through shared memory (rows padded to 33 words) so that every global access is coalesced.
#include "common.cuh"

#include <new>

namespace {

constexpr int kChannelsPerCta = 32;
constexpr int kTileSamples = 64;
constexpr int kRowPitch = kTileSamples + 2;   // int16 elements: 33 words per row -> conflict-free column walks

struct DvfState {
    float xv[11];
    float yv[11];
};

struct DvfParams {
    const int16_t* in;
    int16_t* out;
    unsigned long long in_pitch, out_pitch;
    DvfState* state;
    int channels;
    int n;
};

// (short) of a float the way the x86-64 build does it: cvttss2si (integer indefinite 0x80000000 when out of
// range or NaN), then the low 16 bits
__device__ __forceinline__ int16_t to_short_x86(float v) {
    int r;
    if (v > -2147483904.0f && v < 2147483648.0f) r = __float2int_rz(v);
    else r = (int) 0x80000000u;
    return (int16_t) (r & 0xFFFF);
}

__global__ void __launch_bounds__(kChannelsPerCta * 4) dvf_kernel(const __grid_constant__ DvfParams p) {
    __shared__ int16_t tile[kChannelsPerCta][kRowPitch];
    const int ch0 = blockIdx.x * kChannelsPerCta;
    const int tid = threadIdx.x;
    const int nch = min(kChannelsPerCta, p.channels - ch0);

    // Butterworth band-pass 200..3400 Hz @ 8 kHz, order 5 (mkfilter -Bu -Bp -o 5 -a 0.025 0.425): feedback taps
    const double c0 = 0.1254306222, c1 = 0.1285714097, c2 = -0.8106454980, c3 = -0.7664515771, c4 = 2.1846187758,
                 c5 = 1.8106678608, c6 = -3.1465011600, c7 = -2.0391991609, c8 = 2.4873968618, c9 = 1.0249072542;

    float xv[11], yv[11];
    const bool worker = tid < nch;   // the first warp runs the recurrences, all four warps move data
    if (worker) {
        const DvfState s = p.state[ch0 + tid];
#pragma unroll
        for (int i = 0; i < 11; i++) {
            xv[i] = s.xv[i];
            yv[i] = s.yv[i];
        }
    }

    for (int t0 = 0; t0 < p.n; t0 += kTileSamples) {
        const int len = min(kTileSamples, p.n - t0);
        // coalesced load: 128 threads cover 32 rows x 64 samples in 16 passes (4 rows x 32 lanes x 1 sample ...)
        for (int e = tid; e < kChannelsPerCta * kTileSamples; e += blockDim.x) {
            const int r = e / kTileSamples, c = e % kTileSamples;
            if (r < nch && c < len) tile[r][c] = p.in[(size_t) (ch0 + r) * p.in_pitch + t0 + c];
        }
        __syncthreads();
        if (worker) {
            for (int i = 0; i < len; i++) {
                const float x = __fdiv_rn((float) tile[tid][i], 32767.0f);
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    xv[k] = xv[k + 1];
                    yv[k] = yv[k + 1];
                }
                xv[10] = __fdiv_rn(x, 5.0f);
                float f = __fsub_rn(xv[10], xv[0]);
                f = __fadd_rn(f, __fmul_rn(5.0f, __fsub_rn(xv[2], xv[8])));
                f = __fadd_rn(f, __fmul_rn(10.0f, __fsub_rn(xv[6], xv[4])));
                double acc = (double) f;
                acc = __dadd_rn(acc, __dmul_rn(c0, (double) yv[0]));
                acc = __dadd_rn(acc, __dmul_rn(c1, (double) yv[1]));
                acc = __dadd_rn(acc, __dmul_rn(c2, (double) yv[2]));
                acc = __dadd_rn(acc, __dmul_rn(c3, (double) yv[3]));
                acc = __dadd_rn(acc, __dmul_rn(c4, (double) yv[4]));
                acc = __dadd_rn(acc, __dmul_rn(c5, (double) yv[5]));
                acc = __dadd_rn(acc, __dmul_rn(c6, (double) yv[6]));
                acc = __dadd_rn(acc, __dmul_rn(c7, (double) yv[7]));
                acc = __dadd_rn(acc, __dmul_rn(c8, (double) yv[8]));
                acc = __dadd_rn(acc, __dmul_rn(c9, (double) yv[9]));
                yv[10] = __double2float_rn(acc);
                tile[tid][i] = to_short_x86(__fmul_rn(yv[10], 32767.0f));
            }
        }
        __syncthreads();
        for (int e = tid; e < kChannelsPerCta * kTileSamples; e += blockDim.x) {
            const int r = e / kTileSamples, c = e % kTileSamples;
            if (r < nch && c < len) p.out[(size_t) (ch0 + r) * p.out_pitch + t0 + c] = tile[r][c];
        }
        __syncthreads();
    }

    if (worker) {
        DvfState s;
#pragma unroll
        for (int i = 0; i < 11; i++) {
            s.xv[i] = xv[i];
            s.yv[i] = yv[i];
        }
        p.state[ch0 + tid] = s;
    }
}

}  // namespace

struct dh_dvf {
    int device = 0;
    uint32_t channels = 0;
    DvfState* d_state = nullptr;
};

extern "C" {

int dh_dvf_create(dh_dvf** out, int device, uint32_t channels) {
    DH_REQUIRE(out != nullptr, DH_E_INVALID, "dh_dvf_create: out is NULL");
    *out = nullptr;
    DH_REQUIRE(channels > 0, DH_E_INVALID, "dh_dvf_create: channels must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        dh::set_error("dh_dvf_create: no CUDA device available (this library has no CPU fallback)");
        return DH_E_NODEVICE;
    }
    DH_REQUIRE(device >= 0 && device < ndev, DH_E_INVALID, "dh_dvf_create: device %d out of range", device);
    dh::DeviceGuard guard(device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", device);
    dh_dvf* h = new (std::nothrow) dh_dvf();
    DH_REQUIRE(h != nullptr, DH_E_NOMEM, "dh_dvf_create: out of host memory");
    h->device = device;
    h->channels = channels;
    cudaError_t e = cudaMalloc(&h->d_state, (size_t) channels * sizeof(DvfState));
    if (e == cudaSuccess) e = cudaMemset(h->d_state, 0, (size_t) channels * sizeof(DvfState));
    if (e != cudaSuccess) {
        dh::set_error("dh_dvf_create: %s", cudaGetErrorString(e));
        cudaFree(h->d_state);
        delete h;
        return (int) e;
    }
    *out = h;
    return DH_OK;
}

int dh_dvf_process(dh_dvf* h, const int16_t* d_in, size_t in_pitch, int16_t* d_out, size_t out_pitch, size_t n,
                   void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_dvf_process: handle is NULL");
    if (n == 0) return DH_OK;
    DH_REQUIRE(d_in != nullptr && d_out != nullptr, DH_E_INVALID, "dh_dvf_process: NULL buffer");
    DH_REQUIRE(in_pitch >= n && out_pitch >= n, DH_E_INVALID, "dh_dvf_process: pitch < n");
    DH_REQUIRE(n <= 0x7fffffffu, DH_E_INVALID, "dh_dvf_process: n too large");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    DvfParams p;
    p.in = d_in;
    p.out = d_out;
    p.in_pitch = in_pitch;
    p.out_pitch = out_pitch;
    p.state = h->d_state;
    p.channels = (int) h->channels;
    p.n = (int) n;
    const unsigned grid = (h->channels + kChannelsPerCta - 1) / kChannelsPerCta;
    dvf_kernel<<<grid, kChannelsPerCta * 4, 0, (cudaStream_t) stream>>>(p);
    DH_CUDA(cudaGetLastError());
    return DH_OK;
}

int dh_dvf_reset(dh_dvf* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_dvf_reset: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    DH_CUDA(cudaMemsetAsync(h->d_state, 0, (size_t) h->channels * sizeof(DvfState), (cudaStream_t) stream));
    return DH_OK;
}

uint32_t dh_dvf_channels(const dh_dvf* h) { return h ? h->channels : 0; }

void dh_dvf_destroy(dh_dvf* h) {
    if (!h) return;
    dh::DeviceGuard guard(h->device);
    cudaFree(h->d_state);
    delete h;
}

}  // extern "C"
