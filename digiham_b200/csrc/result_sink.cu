// result_sink.cu — host side of the decoder results: device result blocks -> per-channel byte streams + metadata lines.
//
// The decoder kernels leave fixed-slot rows (`[channels][cap]` bytes / 16-byte event records) plus per-channel counts
// in HBM.  A sink copies only the used widths to pinned memory, appends the bytes and replays the events through the
// per-channel metadata collectors (meta_replay.cu), which reproduce what Digiham::MetaCollector + FileMetaWriter +
// StringSerializer write in the reference (src/lib/meta.cpp:8-17,42-46,58-100).  Used by the decoder bank
// (decoder.cu) for its own channels and by the gathering rank of a sharded pipe (shard.cu) for every rank's.
#include "decoder_ops.hpp"

#include <algorithm>
#include <thread>
#include <vector>

namespace dh {

namespace {

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
    DH_REQUIRE((size_t) c0 + n <= channels, DH_E_INVALID, "result sink: channel range out of bounds");
    if (n == 0) return DH_OK;
    DH_CUDA(cudaMemcpyAsync(h_counts, d_counts, 3 * (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DH_CUDA(cudaStreamSynchronize(st));
    total_d2h += 3 * (uint64_t) n * sizeof(uint32_t);
    const uint32_t* out_len = h_counts;
    const uint32_t* ev_len = h_counts + n;
    const uint32_t* flags = h_counts + 2 * (size_t) n;
    uint32_t max_out = 0, max_ev = 0, any_flags = 0;
    for (uint32_t c = 0; c < n; c++) {
        max_out = std::max(max_out, out_len[c]);
        max_ev = std::max(max_ev, ev_len[c]);
        any_flags |= flags[c];
    }
    if (flags_out) *flags_out |= any_flags;
    DH_REQUIRE(max_out <= out_pitch && max_ev <= ev_pitch, DH_E_STATE,
               "result sink: device counts exceed the slot widths (%u > %zu or %u > %zu)", max_out, out_pitch, max_ev,
               ev_pitch);
    if (max_out) {
        int rc = grow_pinned((void**) &h_out, &h_out_bytes, (size_t) n * max_out);
        if (rc != DH_OK) return rc;
        DH_CUDA(cudaMemcpy2DAsync(h_out, max_out, d_out, out_pitch, max_out, n, cudaMemcpyDeviceToHost, st));
        total_d2h += (uint64_t) n * max_out;
    }
    if (max_ev) {
        const size_t w = (size_t) max_ev * sizeof(DecEvent);
        int rc = grow_pinned((void**) &h_ev, &h_ev_bytes, (size_t) n * w);
        if (rc != DH_OK) return rc;
        DH_CUDA(cudaMemcpy2DAsync(h_ev, w, d_ev, ev_pitch * sizeof(DecEvent), w, n, cudaMemcpyDeviceToHost, st));
        total_d2h += (uint64_t) n * w;
    }
    DH_CUDA(cudaStreamSynchronize(st));
    // per-channel appends and metadata replay are independent: spread them over a few host threads
    auto work = [&](uint32_t a, uint32_t b, uint64_t* sums) {
        for (uint32_t c = a; c < b; c++) {
            ChannelResult& r = results[c0 + c];
            if (out_len[c]) {
                r.bytes.append(reinterpret_cast<const char*>(h_out + (size_t) c * max_out), out_len[c]);
                sums[0] += out_len[c];
            }
            sums[2] += ev_len[c];
            MetaReplay* rp = replay[c0 + c];
            if (ev_len[c] && rp) {
                const size_t before = r.meta.size();
                rp->kv_sink = &r.meta_kv;
                rp->apply_with_output(h_ev + (size_t) c * max_ev, ev_len[c], h_out + (size_t) c * max_out, out_len[c], r.meta);
                sums[1] += r.meta.size() - before;
            }
        }
    };
    unsigned nthreads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
    if (n < 256) nthreads = 1;
    std::vector<uint64_t> sums((size_t) nthreads * 8, 0);   // 8 slots apart: no false sharing
    if (nthreads == 1) {
        work(0, n, sums.data());
    } else {
        std::vector<std::thread> pool;
        const uint32_t per = (n + nthreads - 1) / nthreads;
        for (unsigned t = 0; t < nthreads; t++) {
            const uint32_t a = std::min(n, t * per), b = std::min(n, (t + 1) * per);
            pool.emplace_back(work, a, b, sums.data() + (size_t) t * 8);
        }
        for (auto& t : pool) t.join();
    }
    for (unsigned t = 0; t < nthreads; t++) {
        total_bytes += sums[(size_t) t * 8];
        total_meta += sums[(size_t) t * 8 + 1];
        total_events += sums[(size_t) t * 8 + 2];
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
    if (h_counts) cudaFreeHost(h_counts);
    if (h_out) cudaFreeHost(h_out);
    if (h_ev) cudaFreeHost(h_ev);
    h_counts = nullptr;
    h_out = nullptr;
    h_ev = nullptr;
    h_out_bytes = h_ev_bytes = 0;
    for (auto* r : replay) delete r;
    replay.clear();
    results.clear();
}

}  // namespace dh
