// result_sink.cu — host side of the decoder results: device result blocks -> per-channel byte streams + metadata lines.
//
// The decoder kernels leave fixed-slot rows (`[channels][cap]` bytes / 16-byte event records) plus per-channel counts
// in HBM.  A sink copies only the used widths to pinned memory, appends the bytes and replays the events through the
// per-channel metadata collectors (meta_replay.cu), which reproduce what Digiham::MetaCollector + FileMetaWriter +
// StringSerializer write in the reference (src/lib/meta.cpp:8-17,42-46,58-100).  Used by the decoder bank
// (decoder.cu) for its own channels and by the gathering rank of a sharded pipe (shard.cu) for every rank's.
#include "decoder_ops.hpp"

#include <emmintrin.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dh {

// Persistent worker threads for the replay pass.  They are created once per sink: spawning threads per collect means
// eight stack mmap / munmap pairs and fresh malloc arenas every step, and that address-space churn (TLB shoot-downs,
// IOMMU invalidations under the pinned mappings) was measured to slow the concurrent host -> device upload of the
// next step by 5 % (7.08 -> 7.45 ms per 393 MB block).
class WorkerPool {
    public:
        explicit WorkerPool(unsigned n): count(n) {
            for (unsigned t = 1; t < n; t++) threads.emplace_back([this, t] { loop(t); });
        }
        ~WorkerPool() {
            {
                std::lock_guard<std::mutex> lock(m);
                quit = true;
                generation++;
            }
            wake.notify_all();
            for (auto& t : threads) t.join();
        }
        unsigned size() const { return count; }
        // runs fn(t) for t in [0, size()): t = 0 on the calling thread, the others on the workers; returns when all are done
        void parallel(const std::function<void(unsigned)>& fn) {
            {
                std::lock_guard<std::mutex> lock(m);
                job = &fn;
                pending = count - 1;
                generation++;
            }
            wake.notify_all();
            fn(0);
            std::unique_lock<std::mutex> lock(m);
            done.wait(lock, [this] { return pending == 0; });
            job = nullptr;
        }
    private:
        void loop(unsigned t) {
            uint64_t seen = 0;
            for (;;) {
                const std::function<void(unsigned)>* fn = nullptr;
                {
                    std::unique_lock<std::mutex> lock(m);
                    wake.wait(lock, [&] { return generation != seen; });
                    seen = generation;
                    if (quit) return;
                    fn = job;
                }
                (*fn)(t);
                {
                    std::lock_guard<std::mutex> lock(m);
                    if (--pending == 0) done.notify_one();
                }
            }
        }
        unsigned count;
        std::vector<std::thread> threads;
        std::mutex m;
        std::condition_variable wake, done;
        const std::function<void(unsigned)>* job = nullptr;
        unsigned pending = 0;
        uint64_t generation = 0;
        bool quit = false;
};

namespace {

// dst[r][0 .. w16) = src[r][0 .. w16) for r < rows, in 16-byte units (both pitches are multiples of 16)
__global__ void compact_rows_kernel(const uint4* __restrict__ src, size_t src_pitch16, uint4* __restrict__ dst, uint32_t w16,
                                    uint32_t rows) {
    const size_t total = (size_t) rows * w16;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const size_t r = i / w16, c = i - r * w16;
        dst[i] = src[r * src_pitch16 + c];
    }
}

// Evicts [p, p + n) from the CPU caches.  The staging buffers are written by the GPU's copy engine every step; if
// the lines the replay has just read are still cached, every one of those inbound PCIe writes has to invalidate a
// line in a CPU cache first, and that stalls the root complex long enough to slow the host -> device upload that
// runs at the same time: measured 7.08 -> 7.40 ms per 393 MB block, with nothing but 3 MB of reads as the cause
// (tools/pipe_trace_run.py).  Flushing after use keeps the next step's writes snoop-free.
inline void flush_lines(const void* p, size_t n) {
    if (n == 0) return;
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t) 63;
    const uintptr_t hi = reinterpret_cast<uintptr_t>(p) + n;
    for (uintptr_t a = lo; a < hi; a += 64) _mm_clflush(reinterpret_cast<const void*>(a));
}

int grow_device(uint8_t** p, size_t* have, size_t need) {
    if (need <= *have) return DH_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    need = need + need / 2 + 4096;
    DH_CUDA(cudaMalloc((void**) p, need));
    *have = need;
    return DH_OK;
}

int grow_pinned(void** p, size_t* have, size_t need) {
    if (need <= *have) return DH_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    need = need + need / 2 + 4096;
    DH_CUDA(cudaHostAlloc(p, need, cudaHostAllocDefault));
    *have = need;
    return DH_OK;
}

}  // namespace

int ResultSink::init(int proto, uint32_t nchannels) {
    const ProtoOps* ops = proto_ops(proto);
    DH_REQUIRE(ops != nullptr, DH_E_UNSUPPORTED, "result sink: protocol %d not supported", proto);
    channels = nchannels;
    want_kv = nchannels <= 64;
    results.resize(nchannels);
    replay.assign(nchannels, nullptr);
    if (ops->make_replay) {
        for (uint32_t c = 0; c < nchannels; c++) replay[c] = ops->make_replay();
    }
    DH_CUDA(cudaHostAlloc((void**) &h_counts, 3 * (size_t) nchannels * sizeof(uint32_t), cudaHostAllocDefault));
    return DH_OK;
}

int ResultSink::ingest(const uint32_t* d_counts, const uint8_t* d_out, size_t out_pitch, const DecEvent* d_ev,
                       size_t ev_pitch, uint32_t n, uint32_t c0, cudaStream_t st, uint32_t* flags_out) {
    const Block b = {d_counts, d_out, d_ev, n, c0};
    return ingest_blocks(&b, 1, out_pitch, ev_pitch, st, flags_out);
}

int ResultSink::ingest_blocks(const Block* blocks, int nblocks, size_t out_pitch, size_t ev_pitch, cudaStream_t st,
                              uint32_t* flags_out) {
    struct Plan {
        size_t counts_off;   // into h_counts (uint32 units)
        size_t out_off;      // into h_out (bytes)
        size_t ev_off;       // into h_ev (records)
        uint32_t max_out, max_ev;
    };
    std::vector<Plan> plan((size_t) nblocks);
    size_t total = 0;
    for (int b = 0; b < nblocks; b++) {
        DH_REQUIRE((size_t) blocks[b].c0 + blocks[b].n <= channels, DH_E_INVALID, "result sink: channel range out of bounds");
        plan[b].counts_off = 3 * total;
        total += blocks[b].n;
    }
    DH_REQUIRE(total <= channels, DH_E_INVALID, "result sink: more channels than the sink holds");
    if (total == 0) return DH_OK;
    static const bool timing = getenv("DH_SINK_TIMING") != nullptr;   // diagnostics: pass times on stderr
    static const bool no_flush = getenv("DH_SINK_NO_FLUSH") != nullptr;   // A/B switch for the cache-line flush below
    const auto t0 = std::chrono::steady_clock::now();
    // pass 1: the per-channel counts of every block
    for (int b = 0; b < nblocks; b++) {
        if (blocks[b].n == 0) continue;
        DH_CUDA(cudaMemcpyAsync(h_counts + plan[b].counts_off, blocks[b].d_counts, 3 * (size_t) blocks[b].n * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, st));
    }
    DH_CUDA(cudaStreamSynchronize(st));
    total_d2h += 3 * (uint64_t) total * sizeof(uint32_t);
    uint32_t any_flags = 0;
    size_t out_bytes = 0, ev_records = 0;
    for (int b = 0; b < nblocks; b++) {
        const uint32_t n = blocks[b].n;
        const uint32_t* out_len = h_counts + plan[b].counts_off;
        const uint32_t* ev_len = out_len + n;
        const uint32_t* flags = out_len + 2 * (size_t) n;
        uint32_t max_out = 0, max_ev = 0;
        for (uint32_t c = 0; c < n; c++) {
            max_out = std::max(max_out, out_len[c]);
            max_ev = std::max(max_ev, ev_len[c]);
            any_flags |= flags[c];
        }
        DH_REQUIRE(max_out <= out_pitch && max_ev <= ev_pitch, DH_E_STATE,
                   "result sink: device counts exceed the slot widths (%u > %zu or %u > %zu)", max_out, out_pitch, max_ev,
                   ev_pitch);
        // row widths in the dense staging: bytes rounded up to 16 (the rows are copied in 16-byte units)
        plan[b].max_out = (max_out + 15u) & ~15u;
        plan[b].max_ev = max_ev;
        plan[b].out_off = out_bytes;
        plan[b].ev_off = ev_records;
        out_bytes += (size_t) n * plan[b].max_out;
        ev_records += (size_t) n * max_ev;
    }
    if (flags_out) *flags_out |= any_flags;
    DH_REQUIRE(out_pitch % 16 == 0, DH_E_INVALID, "result sink: byte rows must have a pitch that is a multiple of 16");
    // pass 2: the used widths of all rows are packed densely on the device (a 2-D copy of thousands of sub-kilobyte
    // rows keeps a copy engine busy for a third of a millisecond and stalls the uploads that share it), then ONE
    // contiguous copy per kind brings them to the host
    const size_t ev_bytes = ev_records * sizeof(DecEvent);
    int rc = grow_pinned((void**) &h_out, &h_out_bytes, out_bytes);
    if (rc == DH_OK) rc = grow_pinned((void**) &h_ev, &h_ev_bytes, ev_bytes);
    if (rc == DH_OK) rc = grow_device(&d_compact, &d_compact_bytes, out_bytes + ev_bytes);
    if (rc != DH_OK) return rc;
    for (int b = 0; b < nblocks; b++) {
        const uint32_t n = blocks[b].n;
        if (plan[b].max_out) {
            const uint32_t w16 = plan[b].max_out / 16;
            const unsigned grid = (unsigned) std::min<size_t>(((size_t) n * w16 + 255) / 256, 148 * 8);
            compact_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(blocks[b].d_out), out_pitch / 16,
                                                     reinterpret_cast<uint4*>(d_compact + plan[b].out_off), w16, n);
        }
        if (plan[b].max_ev) {
            const uint32_t w16 = plan[b].max_ev;   // one event record = 16 bytes
            const unsigned grid = (unsigned) std::min<size_t>(((size_t) n * w16 + 255) / 256, 148 * 8);
            compact_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(blocks[b].d_ev), ev_pitch,
                                                     reinterpret_cast<uint4*>(d_compact + out_bytes + plan[b].ev_off * sizeof(DecEvent)),
                                                     w16, n);
        }
    }
    DH_CUDA(cudaGetLastError());
    if (out_bytes) DH_CUDA(cudaMemcpyAsync(h_out, d_compact, out_bytes, cudaMemcpyDeviceToHost, st));
    if (ev_bytes) DH_CUDA(cudaMemcpyAsync(h_ev, d_compact + out_bytes, ev_bytes, cudaMemcpyDeviceToHost, st));
    DH_CUDA(cudaStreamSynchronize(st));
    const auto t1 = std::chrono::steady_clock::now();
    total_d2h += out_bytes + ev_bytes;
    // pass 3: per-channel appends and metadata replay are independent: one parallel pass over all blocks
    auto work = [&](int b, uint32_t c_lo, uint32_t c_hi, uint64_t* sums) {
        const uint32_t n = blocks[b].n;
        const uint32_t* out_len = h_counts + plan[b].counts_off;
        const uint32_t* ev_len = out_len + n;
        const uint32_t max_out = plan[b].max_out, max_ev = plan[b].max_ev;
        for (uint32_t c = c_lo; c < c_hi; c++) {
            ChannelResult& r = results[blocks[b].c0 + c];
            const uint8_t* bytes = h_out + plan[b].out_off + (size_t) c * max_out;
            if (out_len[c]) {
                r.bytes.append(reinterpret_cast<const char*>(bytes), out_len[c]);
                sums[0] += out_len[c];
            }
            sums[2] += ev_len[c];
            MetaReplay* rp = replay[blocks[b].c0 + c];
            if (ev_len[c] && rp) {
                const size_t before = r.meta.size();
                rp->kv_sink = want_kv ? &r.meta_kv : nullptr;
                rp->apply_with_output(h_ev + plan[b].ev_off + (size_t) c * max_ev, ev_len[c], bytes, out_len[c], r.meta);
                sums[1] += r.meta.size() - before;
            }
        }
        if (!no_flush) {
            flush_lines(h_out + plan[b].out_off + (size_t) c_lo * max_out, (size_t) (c_hi - c_lo) * max_out);
            flush_lines(h_ev + plan[b].ev_off + (size_t) c_lo * max_ev, (size_t) (c_hi - c_lo) * max_ev * sizeof(DecEvent));
            flush_lines(out_len + c_lo, (size_t) (c_hi - c_lo) * sizeof(uint32_t));
            flush_lines(ev_len + c_lo, (size_t) (c_hi - c_lo) * sizeof(uint32_t));
            flush_lines(out_len + 2 * (size_t) n + c_lo, (size_t) (c_hi - c_lo) * sizeof(uint32_t));
        }
    };
    unsigned nthreads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), total >= 32768 ? 32u : (total >= 16384 ? 24u : 8u));
    if (total < 256) nthreads = 1;
    static const int forced_threads = getenv("DH_SINK_THREADS") ? atoi(getenv("DH_SINK_THREADS")) : 0;   // tuning switch
    if (forced_threads > 0) nthreads = (unsigned) forced_threads;
    std::vector<uint64_t> sums((size_t) nthreads * 8, 0);   // 8 slots apart: no false sharing
    // thread t takes the t-th slice of the concatenated channel list
    auto run = [&](unsigned t) {
        const size_t per = (total + nthreads - 1) / nthreads;
        size_t lo = std::min(total, (size_t) t * per), hi = std::min(total, (size_t) (t + 1) * per), base = 0;
        for (int b = 0; b < nblocks && lo < hi; b++) {
            const size_t n = blocks[b].n;
            if (lo < base + n) {
                const size_t a = lo - base, e = std::min(n, hi - base);
                work(b, (uint32_t) a, (uint32_t) e, sums.data() + (size_t) t * 8);
                lo = base + e;
            }
            base += n;
        }
    };
    if (nthreads == 1) {
        run(0);
    } else {
        if (pool && pool->size() != nthreads) {
            delete pool;
            pool = nullptr;
        }
        if (!pool) pool = new WorkerPool(nthreads);
        pool->parallel(run);
    }
    uint64_t ev_now = 0, meta_now = 0;
    for (unsigned t = 0; t < nthreads; t++) {
        total_bytes += sums[(size_t) t * 8];
        total_meta += sums[(size_t) t * 8 + 1];
        total_events += sums[(size_t) t * 8 + 2];
        meta_now += sums[(size_t) t * 8 + 1];
        ev_now += sums[(size_t) t * 8 + 2];
    }
    if (timing) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[sink] %zu channels, %d blocks: copies %.2f ms (%.1f MB), replay %.2f ms on %u threads (%llu events, %llu meta bytes)\n",
                total, nblocks, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                (out_bytes + ev_records * sizeof(DecEvent)) / 1e6, std::chrono::duration<double, std::milli>(t2 - t1).count(),
                nthreads, (unsigned long long) ev_now, (unsigned long long) meta_now);
    }
    return DH_OK;
}

void ResultSink::clear() {
    for (auto& r : results) {
        r.bytes.clear();
        r.meta.clear();
        r.meta_kv.clear();
    }
}

void ResultSink::release() {
    delete pool;
    pool = nullptr;
    if (h_counts) cudaFreeHost(h_counts);
    if (h_out) cudaFreeHost(h_out);
    if (h_ev) cudaFreeHost(h_ev);
    cudaFree(d_compact);
    d_compact = nullptr;
    d_compact_bytes = 0;
    h_counts = nullptr;
    h_out = nullptr;
    h_ev = nullptr;
    h_out_bytes = h_ev_bytes = 0;
    for (auto* r : replay) delete r;
    replay.clear();
    results.clear();
}

}  // namespace dh
