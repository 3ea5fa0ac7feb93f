// sharded_pipe.cpp — a C++ host driving one DMR pipe over several GPUs through the C ABI alone:
// one process per GPU, no Python, no torch; NCCL is only reached through dh_shard_*.
//
//   g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include examples/sharded_pipe.cpp -Ldigiham_b200 -ldigiham_b200
//       -Wl,-rpath,$PWD/digiham_b200 -L/usr/local/cuda/lib64 -lcudart -o sharded_pipe          (one command line)
//   ./sharded_pipe <rank> <world> <id file> <input.s16> <channels> <samples per step> <steps> <output prefix>
//
// Start the same command once per rank (rank = GPU index).  Rank 0 is the ingest rank: it reads
// [steps][channels][samples] int16 discriminator samples (what `rtl_fm -M fm -s 48000` emits, reference
// examples/dmr-decoder.sh:12), scatters every step's block, and ends up with the decoded voice frames and metadata
// lines of ALL channels, which it writes as <prefix>.out / <prefix>.meta (per channel: u32 length + bytes).
// The 128-byte NCCL id travels through a file here; any transport the host has will do.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "digiham_b200.h"

#define CHECK(call)                                                                             \
    do {                                                                                        \
        int rc__ = (call);                                                                      \
        if (rc__ != DH_OK) {                                                                    \
            std::fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc__, dh_last_error()); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 9) {
        std::fprintf(stderr, "usage: %s rank world id_file input.s16 channels samples steps out_prefix\n", argv[0]);
        return 2;
    }
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
    const std::string id_file = argv[3], in_file = argv[4], prefix = argv[8];
    const uint64_t channels = std::strtoull(argv[5], nullptr, 10);
    const size_t n = std::strtoull(argv[6], nullptr, 10);
    const int steps = std::atoi(argv[7]);
    if (cudaSetDevice(rank) != cudaSuccess) return 1;

    // bootstrap: rank 0 creates the id, everybody else waits for the file
    void* comm = nullptr;
    if (world > 1) {
        uint8_t id[128];
        if (rank == 0) {
            CHECK(dh_shard_unique_id(id));
            const std::string tmp = id_file + ".tmp";
            FILE* f = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) return 1;
            std::fclose(f);
            std::rename(tmp.c_str(), id_file.c_str());
        } else {
            FILE* f = nullptr;
            for (int tries = 0; tries < 3000 && !(f = std::fopen(id_file.c_str(), "rb")); tries++) usleep(10000);
            if (!f || std::fread(id, 1, sizeof(id), f) != sizeof(id)) return 1;
            std::fclose(f);
        }
        CHECK(dh_shard_comm_init(&comm, id, rank, world, rank));
    }

    dh_shard* sh = nullptr;
    CHECK(dh_shard_create(&sh, comm, rank, world, /*root*/ 0, /*device*/ rank, channels, DH_PROTO_DMR, n, DH_FMT_S16));
    const size_t pitch = dh_shard_pitch(sh);

    // the ingest rank keeps two device blocks [channels][pitch] and refills one while the other is in flight
    int16_t* d_block[2] = {nullptr, nullptr};
    std::vector<int16_t> host;
    FILE* in = nullptr;
    cudaStream_t stream = nullptr;
    if (cudaStreamCreate(&stream) != cudaSuccess) return 1;
    if (rank == 0) {
        in = std::fopen(in_file.c_str(), "rb");
        if (!in) return 1;
        host.resize((size_t) channels * n);
        for (int b = 0; b < 2; b++) {
            if (cudaMalloc(&d_block[b], (size_t) channels * pitch * sizeof(int16_t)) != cudaSuccess) return 1;
            cudaMemset(d_block[b], 0, (size_t) channels * pitch * sizeof(int16_t));
        }
    }
    for (int k = 0; k < steps; k++) {
        if (rank == 0) {
            if (std::fread(host.data(), sizeof(int16_t), host.size(), in) != host.size()) return 1;
            // the block used two steps ago is free once that step has been collected (below)
            if (cudaMemcpy2DAsync(d_block[k & 1], pitch * sizeof(int16_t), host.data(), n * sizeof(int16_t), n * sizeof(int16_t),
                                  channels, cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
            cudaStreamSynchronize(stream);   // `host` is pageable and reused
        }
        CHECK(dh_shard_submit_device(sh, rank == 0 ? d_block[k & 1] : nullptr, pitch, n, DH_SHARD_SCATTER, stream));
        if (k > 0) CHECK(dh_shard_collect_step(sh));   // step k - 1
    }
    CHECK(dh_shard_collect_step(sh));
    CHECK(dh_shard_sync(sh, stream));
    cudaStreamSynchronize(stream);

    if (rank == 0) {
        FILE* fo = std::fopen((prefix + ".out").c_str(), "wb");
        FILE* fm = std::fopen((prefix + ".meta").c_str(), "wb");
        if (!fo || !fm) return 1;
        uint64_t total = 0;
        for (uint64_t c = 0; c < channels; c++) {
            const uint8_t* data = nullptr;
            const char* text = nullptr;
            size_t len = 0, mlen = 0;
            CHECK(dh_shard_output(sh, c, &data, &len));
            CHECK(dh_shard_meta(sh, c, &text, &mlen));
            const uint32_t l32 = (uint32_t) len, m32 = (uint32_t) mlen;
            std::fwrite(&l32, 4, 1, fo);
            std::fwrite(data, 1, len, fo);
            std::fwrite(&m32, 4, 1, fm);
            std::fwrite(text, 1, mlen, fm);
            total += len;
        }
        std::fclose(fo);
        std::fclose(fm);
        uint64_t launches = 0, wire = 0, d2h = 0;
        CHECK(dh_shard_stats(sh, &launches, &wire, &d2h));
        std::printf("%llu channels over %d rank(s) x %d steps: %llu voice bytes gathered on rank 0 (wire block %llu bytes per step)\n",
                    (unsigned long long) channels, world, steps, (unsigned long long) total, (unsigned long long) wire);
    }
    dh_shard_destroy(sh);
    if (comm) dh_shard_comm_destroy(comm);
    cudaFree(d_block[0]);
    cudaFree(d_block[1]);
    return 0;
}
