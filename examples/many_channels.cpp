// many_channels.cpp — the deployment libdigiham_b200 is built for: one bank per GPU, thousands of
// channels, host blocks streamed through the C ABI with two steps in flight.
//
//   g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include examples/many_channels.cpp -Ldigiham_b200 -ldigiham_b200
//       -Wl,-rpath,$PWD/digiham_b200 -L/usr/local/cuda/lib64 -lcudart -o many_channels      (one command line)
//   ./many_channels [channels] [steps]
//
// Every channel carries the same synthetic 4-level signal here (the point is the call sequence, the tests and
// bench.py use real DMR traffic): rrc_filter | gfsk_demodulator | dmr_decoder of examples/dmr-decoder.sh, N times.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "digiham_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc__ = (call);                                                           \
        if (rc__ != DH_OK) {                                                         \
            std::fprintf(stderr, "%s -> %d: %s\n", #call, rc__, dh_last_error());    \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    const uint32_t channels = argc > 1 ? (uint32_t) std::atoi(argv[1]) : 1024;
    const int steps = argc > 2 ? std::atoi(argv[2]) : 8;
    const size_t n = 48000;   // one second of 48 kHz discriminator audio per channel and step

    dh_pipe* pipe = nullptr;
    CHECK(dh_pipe_create(&pipe, /*device*/ 0, channels, DH_PROTO_DMR, n));
    const size_t pitch = dh_pipe_host_pitch(pipe);

    // two pinned blocks [channels][pitch]: the upload of step k + 1 overlaps the kernels and the read-back of step k
    float* block[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; b++) {
        if (cudaHostAlloc((void**) &block[b], (size_t) channels * pitch * sizeof(float), cudaHostAllocDefault) != cudaSuccess) return 1;
        for (uint32_t c = 0; c < channels; c++) {
            for (size_t t = 0; t < n; t++) {
                const int symbol = (int) ((t / 10 + 7 * c) % 4);          // 4800 symbols/s, 10 samples per symbol
                block[b][(size_t) c * pitch + t] = 0.5f * (float) (2 * symbol - 3) / 3.0f;
            }
        }
    }

    const auto t0 = std::chrono::steady_clock::now();
    CHECK(dh_pipe_submit_host(pipe, block[0], pitch, n));
    for (int k = 1; k < steps; k++) {
        CHECK(dh_pipe_submit_host(pipe, block[k & 1], pitch, n));   // returns at once
        CHECK(dh_pipe_collect_step(pipe));                          // results of step k - 1
    }
    CHECK(dh_pipe_collect_step(pipe));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // per-channel results: decoded voice frames (27 bytes each) and the metadata lines a FileMetaWriter would write
    dh_decoder* dec = dh_pipe_decoder(pipe);
    uint64_t out_bytes = 0, meta_bytes = 0;
    CHECK(dh_decoder_totals(dec, &out_bytes, &meta_bytes));
    const uint8_t* data = nullptr;
    size_t len = 0;
    CHECK(dh_decoder_output(dec, 0, &data, &len));
    std::printf("%u channels x %d steps: %.1f Msamples/s end to end, %llu voice bytes, %llu metadata bytes (channel 0: %zu bytes)\n",
                channels, steps, (double) channels * n * steps / sec / 1e6, (unsigned long long) out_bytes,
                (unsigned long long) meta_bytes, len);

    dh_pipe_destroy(pipe);
    cudaFreeHost(block[0]);
    cudaFreeHost(block[1]);
    return 0;
}
