// decoder.cu — generic protocol decoder bank (dh_decoder_*): device buffers, zero-copy symbol hand-off, result
// collection and host-side metadata replay.  The per-protocol kernels live in dmr.cu / ysf.cu / pocsag.cu.
//
// Replaces Digiham::Decoder (reference include/decoder.hpp:17-30, src/lib/decoder.cpp:7-47) for N channels.
#include "decoder_ops.hpp"

#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

using dh::DecEvent;
using dh::DecIo;

namespace {
constexpr uint32_t kAccumulate = 2;   // process calls whose results fit between two collects
}

struct dh_decoder {
    int device = 0;
    uint32_t channels = 0;
    int proto = 0;
    const dh::ProtoOps* ops = nullptr;

    void* d_states = nullptr;
    uint8_t* d_sym = nullptr;
    size_t sym_pitch = 0;
    size_t max_syms = 0;
    // two result sets so that a streaming caller can read the results of step k while step k+1 is being decoded
    uint8_t* d_out_set[2] = {nullptr, nullptr};
    uint32_t out_cap = 0;
    DecEvent* d_ev_set[2] = {nullptr, nullptr};
    uint32_t ev_cap = 0;
    uint32_t* d_counts_set[2] = {nullptr, nullptr};   // each [3][channels]: out_len, ev_len, flags
    int active = 0;                    // set written by dh_decoder_process
    uint8_t* d_slot_filter = nullptr;
    std::vector<uint8_t> h_slot_filter;
    bool filter_dirty = true;

    dh::ResultSink sink;               // host side: per-channel results + metadata replay
};

namespace {

int decoder_collect(dh_decoder* h, int set, cudaStream_t st);

int decoder_reserve(dh_decoder* h, size_t max_syms) {
    if (h->d_sym && max_syms <= h->max_syms) return DH_OK;
    const size_t m16 = (max_syms + 15) & ~(size_t) 15;
    const size_t pitch = (size_t) h->ops->carry_cap + m16;
    if (h->d_sym) {
        // mid-stream growth: drain pending results, keep the carried symbol tails
        DH_CUDA(cudaDeviceSynchronize());
        for (int set = 0; set < 2; set++) {
            int rc = decoder_collect(h, set, nullptr);
            if (rc != DH_OK) return rc;
        }
    }
    uint8_t* ns = nullptr;
    DH_CUDA(cudaMalloc(&ns, (size_t) h->channels * pitch));
    DH_CUDA(cudaMemset(ns, 0, (size_t) h->channels * pitch));
    if (h->d_sym) {
        DH_CUDA(cudaMemcpy2D(ns, pitch, h->d_sym, h->sym_pitch, (size_t) h->ops->carry_cap, h->channels,
                             cudaMemcpyDeviceToDevice));
        DH_CUDA(cudaFree(h->d_sym));
        for (int set = 0; set < 2; set++) {
            DH_CUDA(cudaFree(h->d_out_set[set]));
            DH_CUDA(cudaFree(h->d_ev_set[set]));
            h->d_out_set[set] = nullptr;
            h->d_ev_set[set] = nullptr;
        }
    }
    h->d_sym = ns;
    h->sym_pitch = pitch;
    h->max_syms = m16;
    h->out_cap = (h->ops->out_bytes(m16) * kAccumulate + 15u) & ~15u;
    h->ev_cap = h->ops->events(m16) * kAccumulate;
    for (int set = 0; set < 2; set++) {
        DH_CUDA(cudaMalloc(&h->d_out_set[set], (size_t) h->channels * h->out_cap));
        DH_CUDA(cudaMalloc(&h->d_ev_set[set], (size_t) h->channels * h->ev_cap * sizeof(DecEvent)));
        // collect copies the widest channel's byte count for every channel: keep the unused tails defined
        DH_CUDA(cudaMemset(h->d_out_set[set], 0, (size_t) h->channels * h->out_cap));
        DH_CUDA(cudaMemset(h->d_ev_set[set], 0, (size_t) h->channels * h->ev_cap * sizeof(DecEvent)));
    }
    return DH_OK;
}

int decoder_collect(dh_decoder* h, int set, cudaStream_t st) {
    if (!h->d_out_set[set]) return DH_OK;
    uint32_t* d_counts = h->d_counts_set[set];
    uint32_t any_flags = 0;
    int rc = h->sink.ingest(d_counts, h->d_out_set[set], h->out_cap, h->d_ev_set[set], h->ev_cap, h->channels, 0, st,
                            &any_flags);
    if (rc != DH_OK) return rc;
    DH_CUDA(cudaMemsetAsync(d_counts, 0, 3 * (size_t) h->channels * sizeof(uint32_t), st));
    DH_CUDA(cudaStreamSynchronize(st));
    DH_REQUIRE(any_flags == 0, DH_E_STATE,
               "dh_decoder_collect: device result buffers overflowed (flags 0x%x): collect after every %u process calls",
               any_flags, kAccumulate);
    return DH_OK;
}

}  // namespace

namespace dh {

const ProtoOps* proto_ops(int proto) {
    switch (proto) {
        case DH_PROTO_DMR: return dmr_ops();
        case DH_PROTO_POCSAG: return pocsag_ops();
        case DH_PROTO_YSF: return ysf_ops();
        case DH_PROTO_NXDN: return nxdn_ops();
        case DH_PROTO_DSTAR: return dstar_ops();
        default: return nullptr;
    }
}

int decoder_view(dh_decoder* h, int set, DecoderView* view) {
    DH_REQUIRE(h != nullptr && view != nullptr && (set == 0 || set == 1), DH_E_INVALID, "decoder_view: bad argument");
    DH_REQUIRE(h->d_out_set[set] != nullptr, DH_E_STATE, "decoder_view: the bank has no result buffers yet (reserve first)");
    view->out = h->d_out_set[set];
    view->out_cap = h->out_cap;
    view->ev = h->d_ev_set[set];
    view->ev_cap = h->ev_cap;
    view->counts = h->d_counts_set[set];
    view->channels = h->channels;
    view->max_syms = h->max_syms;
    return DH_OK;
}

}  // namespace dh

extern "C" {

int dh_decoder_create(dh_decoder** out, int device, uint32_t channels, int proto) {
    DH_REQUIRE(out != nullptr, DH_E_INVALID, "dh_decoder_create: out is NULL");
    *out = nullptr;
    DH_REQUIRE(channels > 0, DH_E_INVALID, "dh_decoder_create: channels must be > 0");
    const dh::ProtoOps* ops = dh::proto_ops(proto);
    DH_REQUIRE(ops != nullptr, DH_E_UNSUPPORTED, "dh_decoder_create: protocol %d not supported", proto);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        dh::set_error("dh_decoder_create: no CUDA device available (this library has no CPU fallback)");
        return DH_E_NODEVICE;
    }
    DH_REQUIRE(device >= 0 && device < ndev, DH_E_INVALID, "dh_decoder_create: device %d out of range", device);
    dh::DeviceGuard guard(device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", device);
    dh_decoder* h = new (std::nothrow) dh_decoder();
    DH_REQUIRE(h != nullptr, DH_E_NOMEM, "dh_decoder_create: out of host memory");
    h->device = device;
    h->channels = channels;
    h->proto = proto;
    h->ops = ops;
    h->h_slot_filter.assign(channels, 3);
    if (h->sink.init(proto, channels) != DH_OK) {
        delete h;
        return DH_E_NOMEM;
    }
    std::vector<uint8_t> init((size_t) channels * ops->state_size);
    ops->init_states(init.data(), channels);
    cudaError_t e = cudaMalloc(&h->d_states, init.size());
    if (e == cudaSuccess) e = cudaMemcpy(h->d_states, init.data(), init.size(), cudaMemcpyHostToDevice);
    for (int set = 0; set < 2; set++) {
        if (e == cudaSuccess) e = cudaMalloc(&h->d_counts_set[set], 3 * (size_t) channels * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(h->d_counts_set[set], 0, 3 * (size_t) channels * sizeof(uint32_t));
    }
    if (e == cudaSuccess) e = cudaMalloc(&h->d_slot_filter, channels);
    if (e != cudaSuccess) {
        dh::set_error("dh_decoder_create: %s", cudaGetErrorString(e));
        dh_decoder_destroy(h);
        return (int) e;
    }
    *out = h;
    return DH_OK;
}

int dh_decoder_reserve(dh_decoder* h, size_t max_syms, uint8_t** d_buf, size_t* pitch) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_reserve: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    int rc = decoder_reserve(h, max_syms);
    if (rc != DH_OK) return rc;
    if (d_buf) *d_buf = h->d_sym + h->ops->carry_cap;
    if (pitch) *pitch = h->sym_pitch;
    return DH_OK;
}

int dh_decoder_set_slot_filter(dh_decoder* h, int channel, uint8_t filter) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_set_slot_filter: handle is NULL");
    DH_REQUIRE(channel < (int) h->channels, DH_E_INVALID, "dh_decoder_set_slot_filter: channel out of range");
    // the per-channel control byte carries the filter in its low nibble and the decoder options above it
    auto put = [&](uint8_t& b) { b = (uint8_t) ((b & 0xF0u) | (filter & 0x0Fu)); };
    if (channel < 0) for (auto& b : h->h_slot_filter) put(b);
    else put(h->h_slot_filter[channel]);
    h->filter_dirty = true;
    return DH_OK;
}

int dh_decoder_set_option(dh_decoder* h, int channel, int option, int value) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_set_option: handle is NULL");
    DH_REQUIRE(channel < (int) h->channels, DH_E_INVALID, "dh_decoder_set_option: channel out of range");
    DH_REQUIRE(option == DH_OPT_DMR_LC_FEC && h->proto == DH_PROTO_DMR, DH_E_UNSUPPORTED,
               "dh_decoder_set_option: option %d is not available for protocol %d", option, h->proto);
    DH_REQUIRE(value >= 0 && value <= 2, DH_E_INVALID, "dh_decoder_set_option: value %d out of range", value);
    auto put = [&](uint8_t& b) { b = (uint8_t) ((b & 0xCFu) | ((unsigned) value << 4)); };
    if (channel < 0) for (auto& b : h->h_slot_filter) put(b);
    else put(h->h_slot_filter[channel]);
    h->filter_dirty = true;
    return DH_OK;
}

int dh_decoder_process(dh_decoder* h, const uint8_t* d_sym, size_t sym_pitch, const uint32_t* d_nsym, size_t max_nsym,
                       void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_process: handle is NULL");
    DH_REQUIRE(d_nsym != nullptr, DH_E_INVALID, "dh_decoder_process: d_nsym is NULL");
    if (max_nsym == 0) return DH_OK;
    DH_REQUIRE(d_sym != nullptr, DH_E_INVALID, "dh_decoder_process: d_sym is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    cudaStream_t st = (cudaStream_t) stream;
    const bool zero_copy = h->d_sym && d_sym == h->d_sym + h->ops->carry_cap && sym_pitch == h->sym_pitch &&
                           max_nsym <= h->max_syms;
    if (!zero_copy) {
        DH_REQUIRE(sym_pitch >= max_nsym, DH_E_INVALID, "dh_decoder_process: sym_pitch < max_nsym");
        int rc = decoder_reserve(h, max_nsym);
        if (rc != DH_OK) return rc;
        DH_CUDA(cudaMemcpy2DAsync(h->d_sym + h->ops->carry_cap, h->sym_pitch, d_sym, sym_pitch, max_nsym, h->channels,
                                  cudaMemcpyDeviceToDevice, st));
    }
    if (h->filter_dirty) {
        DH_CUDA(cudaMemcpyAsync(h->d_slot_filter, h->h_slot_filter.data(), h->channels, cudaMemcpyHostToDevice, st));
        DH_CUDA(cudaStreamSynchronize(st));   // the source vector may change right after this call
        h->filter_dirty = false;
    }
    DecIo io;
    io.sym = h->d_sym;
    io.sym_pitch = h->sym_pitch;
    io.nsym = d_nsym;
    io.out = h->d_out_set[h->active];
    io.out_len = h->d_counts_set[h->active];
    io.ev = h->d_ev_set[h->active];
    io.ev_len = h->d_counts_set[h->active] + h->channels;
    io.flags = h->d_counts_set[h->active] + 2 * (size_t) h->channels;
    io.out_cap = h->out_cap;
    io.ev_cap = h->ev_cap;
    io.carry_cap = h->ops->carry_cap;
    io.channels = (int) h->channels;
    return h->ops->launch(io, h->d_states, h->d_slot_filter, st);
}

int dh_decoder_collect(dh_decoder* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_collect: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    return decoder_collect(h, h->active, (cudaStream_t) stream);
}

int dh_decoder_select_results(dh_decoder* h, int set) {
    DH_REQUIRE(h != nullptr && (set == 0 || set == 1), DH_E_INVALID, "dh_decoder_select_results: bad argument");
    h->active = set;
    return DH_OK;
}

int dh_decoder_collect_results(dh_decoder* h, int set, void* stream) {
    DH_REQUIRE(h != nullptr && (set == 0 || set == 1), DH_E_INVALID, "dh_decoder_collect_results: bad argument");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    return decoder_collect(h, set, (cudaStream_t) stream);
}

int dh_decoder_output(dh_decoder* h, uint32_t channel, const uint8_t** data, size_t* len) {
    DH_REQUIRE(h != nullptr && channel < h->channels, DH_E_INVALID, "dh_decoder_output: bad handle or channel");
    if (data) *data = reinterpret_cast<const uint8_t*>(h->sink.results[channel].bytes.data());
    if (len) *len = h->sink.results[channel].bytes.size();
    return DH_OK;
}

int dh_decoder_meta(dh_decoder* h, uint32_t channel, const char** text, size_t* len) {
    DH_REQUIRE(h != nullptr && channel < h->channels, DH_E_INVALID, "dh_decoder_meta: bad handle or channel");
    if (text) *text = h->sink.results[channel].meta.data();
    if (len) *len = h->sink.results[channel].meta.size();
    return DH_OK;
}

int dh_decoder_meta_kv(dh_decoder* h, uint32_t channel, const uint8_t** data, size_t* len) {
    DH_REQUIRE(h != nullptr && channel < h->channels, DH_E_INVALID, "dh_decoder_meta_kv: bad handle or channel");
    if (data) *data = reinterpret_cast<const uint8_t*>(h->sink.results[channel].meta_kv.data());
    if (len) *len = h->sink.results[channel].meta_kv.size();
    return DH_OK;
}

int dh_decoder_set_meta_kv(dh_decoder* h, int enable) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_set_meta_kv: handle is NULL");
    h->sink.want_kv = enable != 0;
    return DH_OK;
}

int dh_decoder_totals(dh_decoder* h, uint64_t* out_bytes, uint64_t* meta_bytes) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_totals: handle is NULL");
    if (out_bytes) *out_bytes = h->sink.total_bytes;
    if (meta_bytes) *meta_bytes = h->sink.total_meta;
    return DH_OK;
}

int dh_decoder_stats(dh_decoder* h, uint64_t* events, uint64_t* d2h_bytes) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_stats: handle is NULL");
    if (events) *events = h->sink.total_events;
    if (d2h_bytes) *d2h_bytes = h->sink.total_d2h;
    return DH_OK;
}

int dh_decoder_discard(dh_decoder* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_discard: handle is NULL");
    if (!h->d_counts_set[h->active]) return DH_OK;
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    DH_CUDA(cudaMemsetAsync(h->d_counts_set[h->active], 0, 3 * (size_t) h->channels * sizeof(uint32_t),
                            (cudaStream_t) stream));
    return DH_OK;
}

int dh_decoder_clear(dh_decoder* h) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_clear: handle is NULL");
    h->sink.clear();
    return DH_OK;
}

int dh_meta_replay(int proto, const void* events, uint32_t n_events, char* out, size_t cap, size_t* len) {
    DH_REQUIRE(events != nullptr || n_events == 0, DH_E_INVALID, "dh_meta_replay: events is NULL");
    dh::MetaReplay* r = nullptr;
    switch (proto) {
        case DH_PROTO_DMR: r = dh::make_dmr_replay(); break;
        case DH_PROTO_YSF: r = dh::make_ysf_replay(); break;
        case DH_PROTO_NXDN: r = dh::make_nxdn_replay(); break;
        case DH_PROTO_DSTAR: r = dh::make_dstar_replay(); break;
        default: break;
    }
    DH_REQUIRE(r != nullptr, DH_E_UNSUPPORTED, "dh_meta_replay: protocol %d has no metadata replay", proto);
    std::string text;
    r->apply(static_cast<const DecEvent*>(events), n_events, text);
    delete r;
    if (len) *len = text.size();
    if (out && cap) std::memcpy(out, text.data(), std::min(cap, text.size()));
    return DH_OK;
}

uint32_t dh_decoder_channels(const dh_decoder* h) { return h ? h->channels : 0; }

// ---- state: per-channel phase state + unconsumed symbol tails (device), slot filters and the metadata collectors
// (host).  Results that were decoded but not collected yet are not part of the state: collect first.
static dh::StateHeader decoder_header(const dh_decoder* h, uint64_t payload) {
    return dh::make_state_header(3, h->channels, (uint32_t) h->proto, (uint32_t) h->ops->state_size,
                                 (uint32_t) h->ops->carry_cap, 0, payload);
}

static void decoder_replay_blobs(const dh_decoder* h, std::string& blob) {
    for (uint32_t c = 0; c < h->channels; c++) {
        std::string one;
        if (h->sink.replay[c]) h->sink.replay[c]->save(one);
        const uint32_t n = (uint32_t) one.size();
        blob.append(reinterpret_cast<const char*>(&n), sizeof(n));
        blob.append(one);
    }
}

int dh_decoder_state_size(const dh_decoder* h, size_t* bytes) {
    DH_REQUIRE(h != nullptr && bytes != nullptr, DH_E_INVALID, "dh_decoder_state_size: NULL argument");
    std::string blob;
    decoder_replay_blobs(h, blob);
    *bytes = sizeof(dh::StateHeader) + (size_t) h->channels * (h->ops->state_size + (size_t) h->ops->carry_cap + 1) +
             blob.size();
    return DH_OK;
}

int dh_decoder_state_export(dh_decoder* h, void* h_buf, size_t cap, size_t* written, void* stream) {
    DH_REQUIRE(h != nullptr && h_buf != nullptr, DH_E_INVALID, "dh_decoder_state_export: NULL argument");
    std::string blob;
    decoder_replay_blobs(h, blob);
    const size_t dev_bytes = (size_t) h->channels * (h->ops->state_size + (size_t) h->ops->carry_cap);
    const dh::StateHeader hd = decoder_header(h, dev_bytes + h->channels + blob.size());
    DH_REQUIRE(cap >= sizeof(hd) + hd.payload, DH_E_INVALID, "dh_decoder_state_export: buffer too small");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    int rc = decoder_reserve(h, h->max_syms ? h->max_syms : 16);
    if (rc != DH_OK) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    char* out = static_cast<char*>(h_buf);
    std::memcpy(out, &hd, sizeof(hd));
    out += sizeof(hd);
    DH_CUDA(cudaMemcpyAsync(out, h->d_states, (size_t) h->channels * h->ops->state_size, cudaMemcpyDeviceToHost, st));
    out += (size_t) h->channels * h->ops->state_size;
    DH_CUDA(cudaMemcpy2DAsync(out, h->ops->carry_cap, h->d_sym, h->sym_pitch, h->ops->carry_cap, h->channels,
                              cudaMemcpyDeviceToHost, st));
    out += (size_t) h->channels * h->ops->carry_cap;
    DH_CUDA(cudaStreamSynchronize(st));
    std::memcpy(out, h->h_slot_filter.data(), h->channels);
    out += h->channels;
    std::memcpy(out, blob.data(), blob.size());
    if (written) *written = sizeof(hd) + hd.payload;
    return DH_OK;
}

int dh_decoder_state_import(dh_decoder* h, const void* h_buf, size_t bytes, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_decoder_state_import: handle is NULL");
    const size_t dev_bytes = (size_t) h->channels * (h->ops->state_size + (size_t) h->ops->carry_cap);
    dh::StateHeader hd = decoder_header(h, 0);
    int rc = dh::check_state_header(h_buf, bytes, hd, "dh_decoder_state_import");
    if (rc != DH_OK) return rc;
    std::memcpy(&hd, h_buf, sizeof(hd));
    DH_REQUIRE(hd.payload >= dev_bytes + h->channels, DH_E_INVALID, "dh_decoder_state_import: truncated blob");
    // host part first (it can fail on a malformed blob before anything on the device changes)
    const uint8_t* in = static_cast<const uint8_t*>(h_buf) + sizeof(hd);
    const uint8_t* end = in + hd.payload;
    const uint8_t* p = in + dev_bytes + h->channels;
    std::vector<std::pair<const uint8_t*, uint32_t>> parts(h->channels);
    for (uint32_t c = 0; c < h->channels; c++) {
        uint32_t n = 0;
        DH_REQUIRE((size_t) (end - p) >= sizeof(n), DH_E_INVALID, "dh_decoder_state_import: truncated collector state");
        std::memcpy(&n, p, sizeof(n));
        p += sizeof(n);
        DH_REQUIRE((size_t) (end - p) >= n, DH_E_INVALID, "dh_decoder_state_import: truncated collector state");
        parts[c] = {p, n};
        p += n;
    }
    for (uint32_t c = 0; c < h->channels; c++) {
        if (!h->sink.replay[c]) continue;
        DH_REQUIRE(h->sink.replay[c]->load(parts[c].first, parts[c].second), DH_E_INVALID,
                   "dh_decoder_state_import: malformed collector state of channel %u", c);
    }
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    rc = decoder_reserve(h, h->max_syms ? h->max_syms : 16);
    if (rc != DH_OK) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    DH_CUDA(cudaMemcpyAsync(h->d_states, in, (size_t) h->channels * h->ops->state_size, cudaMemcpyHostToDevice, st));
    in += (size_t) h->channels * h->ops->state_size;
    DH_CUDA(cudaMemcpy2DAsync(h->d_sym, h->sym_pitch, in, h->ops->carry_cap, h->ops->carry_cap, h->channels,
                              cudaMemcpyHostToDevice, st));
    in += (size_t) h->channels * h->ops->carry_cap;
    DH_CUDA(cudaStreamSynchronize(st));
    std::memcpy(h->h_slot_filter.data(), in, h->channels);
    h->filter_dirty = true;
    return DH_OK;
}

void dh_decoder_destroy(dh_decoder* h) {
    if (!h) return;
    dh::DeviceGuard guard(h->device);
    cudaFree(h->d_states);
    cudaFree(h->d_sym);
    for (int set = 0; set < 2; set++) {
        cudaFree(h->d_out_set[set]);
        cudaFree(h->d_ev_set[set]);
        cudaFree(h->d_counts_set[set]);
    }
    cudaFree(h->d_slot_filter);
    h->sink.release();
    delete h;
}

}  // extern "C"
