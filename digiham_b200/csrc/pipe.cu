// pipe.cu — dh_pipe_*: the whole per-channel pipe of the reference's example scripts as one bank object.
//
//   DMR / YSF : WideRrcFilter -> GfskDemodulator(10) -> decoder   (reference examples/dmr-decoder.sh:19-23,
//                                                                  examples/ysf-decoder.sh:19-23)
//   POCSAG    : FskDemodulator(40, invert) -> decoder             (reference examples/pocsag-decoder.sh:19-21)
//   NXDN      : NarrowRrcFilter -> GfskDemodulator(20) -> decoder (reference examples/nxdn48-decoder.sh:19-23)
//   D-Star    : FskDemodulator(10) -> decoder                     (reference examples/dstar-decoder.sh:19-21)
//
// In the reference every `|` is a process boundary with a 1024-item ring in between (src/lib/cli.cpp:10,101-106);
// here the stages are kernels on one stream and each stage writes straight into the next stage's carry-aware
// input rows (dh_demod_reserve / dh_decoder_reserve), so no intermediate copy exists.
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

struct dh_pipe {
    int device = 0;
    uint32_t channels = 0;
    int proto = 0;
    size_t max_chunk = 0;
    dh_rrc* rrc = nullptr;
    dh_demod* demod = nullptr;
    dh_decoder* decoder = nullptr;
    float* d_filt = nullptr;      // demodulator input rows (owned by demod)
    size_t filt_pitch = 0;
    uint8_t* d_sym = nullptr;     // decoder input rows (owned by decoder)
    size_t sym_pitch = 0;
    size_t max_syms = 0;
    uint32_t* d_nsym = nullptr;
    float* d_stage = nullptr;     // device staging for host input (dh_pipe_process_host)
    size_t stage_pitch = 0;
    // streaming host interface (dh_pipe_submit_host / dh_pipe_collect_step): two steps in flight, the upload of
    // step k+1 (copy stream) overlaps kernels + result read-back + metadata replay of step k
    float* d_slot[2] = {nullptr, nullptr};
    cudaStream_t s_copy = nullptr, s_compute = nullptr, s_back = nullptr;
    cudaStream_t s_copy_back() const { return s_back; }
    cudaEvent_t ev_uploaded[2] = {nullptr, nullptr};   // H2D of the slot finished
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};   // K1 finished reading the slot
    cudaEvent_t ev_decoded[2] = {nullptr, nullptr};    // decoder kernel of the step finished
    uint64_t submitted = 0, collected = 0;
    // diagnostics (DH_PIPE_TRACE): timing events around the upload and the kernels of every submitted step
    std::vector<cudaEvent_t> trace;   // groups of 4: upload start / end, kernels start / end
    // optional per-stage device timing: CUDA events around every kernel, on the stream it is launched on
    bool profiling = false;
    std::vector<cudaEvent_t> events;   // pool; groups of 6: K1 start/end, K2 start/end, decoder start/end
    size_t events_used = 0;            // recorded since the last dh_pipe_stage_times
    cudaStream_t last_k1_stream = nullptr;   // stream the first kernel of the most recent call was enqueued on
    uint64_t launches = 0;             // kernels launched by this pipe since creation
    // Software pipelining inside one process call: the chunk is cut into sub-chunks; K1 of sub-chunk c+1 (stream
    // a) overlaps K2 + decoder of sub-chunk c (stream b, higher priority).  K1 is issue-bound, K2 and the decoder
    // are latency-bound sequential walks, so together they fill the SMs.
    size_t sub_chunk = 0;              // 0 = no pipelining (default: measured slower on B200)
    cudaStream_t sa = nullptr, sb = nullptr;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr;
    std::vector<cudaEvent_t> ev_k1, ev_k2;
    // Software pipelining ACROSS process calls (dh_pipe_set_async): K1 of call i+1 (stream a) overlaps K2 + decoder
    // of call i (stream b).  The caller's stream only orders the input; results are joined by dh_pipe_sync /
    // dh_pipe_collect.
    bool async_mode = false;
    uint64_t async_step = 0;
    cudaEvent_t ev_a1[4] = {nullptr, nullptr, nullptr, nullptr}, ev_a2[4] = {nullptr, nullptr, nullptr, nullptr};
    bool async_pending = false;        // work enqueued on the internal streams that `ev_done` covers
    cudaEvent_t ev_last_k1 = nullptr;  // K1 of the most recent asynchronous call (the last reader of its input)
};

extern "C" {

void dh_pipe_destroy(dh_pipe* h);

int dh_pipe_create(dh_pipe** out, int device, uint32_t channels, int proto, size_t max_chunk) {
    DH_REQUIRE(out != nullptr, DH_E_INVALID, "dh_pipe_create: out is NULL");
    *out = nullptr;
    DH_REQUIRE(max_chunk > 0, DH_E_INVALID, "dh_pipe_create: max_chunk must be > 0");
    dh_pipe* h = new (std::nothrow) dh_pipe();
    DH_REQUIRE(h != nullptr, DH_E_NOMEM, "dh_pipe_create: out of host memory");
    h->device = device;
    h->channels = channels;
    h->proto = proto;
    h->max_chunk = max_chunk;
    int rc = DH_OK;
    const bool pocsag = proto == DH_PROTO_POCSAG;
    const bool nxdn = proto == DH_PROTO_NXDN;
    const bool dstar = proto == DH_PROTO_DSTAR;
    if (!pocsag && !dstar) rc = dh_rrc_create(&h->rrc, device, channels, nxdn ? DH_RRC_NARROW : DH_RRC_WIDE);
    if (rc == DH_OK) {
        if (pocsag) rc = dh_demod_create(&h->demod, device, channels, 0, 40, 1);
        else if (dstar) rc = dh_demod_create(&h->demod, device, channels, 0, 10, 0);
        else rc = dh_demod_create(&h->demod, device, channels, 1, nxdn ? 20 : 10, 0);
    }
    if (rc == DH_OK) rc = dh_decoder_create(&h->decoder, device, channels, proto);
    // pipes without an RRC stage hand the caller's block to the demodulator, which reads it in place
    if (rc == DH_OK && h->rrc) rc = dh_demod_reserve(h->demod, max_chunk, &h->d_filt, &h->filt_pitch);
    if (rc == DH_OK) {
        h->max_syms = dh_demod_max_symbols(h->demod, max_chunk);
        rc = dh_decoder_reserve(h->decoder, h->max_syms, &h->d_sym, &h->sym_pitch);
    }
    if (rc == DH_OK) {
        dh::DeviceGuard guard(device);
        DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", device);
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        int pa = lo_prio, pb = hi_prio;
        if (const char* env = getenv("DH_PIPE_PRIO")) {   // experiment switch
            if (atoi(env) == 1) pa = pb = lo_prio;
            if (atoi(env) == 2) { pa = hi_prio; pb = lo_prio; }
        }
        cudaError_t e = cudaStreamCreateWithPriority(&h->sa, cudaStreamNonBlocking, pa);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->sb, cudaStreamNonBlocking, pb);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_nsym, (size_t) channels * sizeof(uint32_t));
        if (e != cudaSuccess) {
            dh::set_error("dh_pipe_create: %s", cudaGetErrorString(e));
            rc = (int) e;
        }
    }
    if (rc != DH_OK) {
        dh_pipe_destroy(h);
        return rc;
    }
    *out = h;
    return DH_OK;
}

namespace {

constexpr size_t kMaxTimedCalls = 65536;

// one sub-chunk through the three stages; k1/k23 are the streams of K1 and of K2 + decoder
int run_stages(dh_pipe* h, const void* d_in, size_t in_pitch, size_t n, cudaStream_t k1, cudaStream_t k23,
               cudaEvent_t k1_done, cudaEvent_t k2_done, bool s16 = false) {
    // per-stage timing events come from a pool that is recycled by dh_pipe_stage_times; a profiled pipe that is
    // never queried stops recording once kMaxTimedCalls calls are pending instead of growing without bound
    auto mark = [&](cudaStream_t st) -> int {
        if (!h->profiling || h->events_used >= kMaxTimedCalls * 6) return DH_OK;
        if (h->events_used == h->events.size()) {
            cudaEvent_t e;
            DH_CUDA(cudaEventCreate(&e));
            h->events.push_back(e);
        }
        DH_CUDA(cudaEventRecord(h->events[h->events_used++], st));
        return DH_OK;
    };
    h->last_k1_stream = k1;
    int rc;
    if (h->rrc) {
        // the demodulator's input rows alternate between two buffers from call to call
        rc = dh_demod_reserve(h->demod, h->max_chunk, &h->d_filt, &h->filt_pitch);
        if (rc != DH_OK) return rc;
        if ((rc = mark(k1)) != DH_OK) return rc;
        rc = s16 ? dh_rrc_process_s16(h->rrc, static_cast<const int16_t*>(d_in), in_pitch, h->d_filt, h->filt_pitch, n, k1)
                 : dh_rrc_process(h->rrc, static_cast<const float*>(d_in), in_pitch, h->d_filt, h->filt_pitch, n, k1);
        if (rc != DH_OK) return rc;
        h->launches++;
        if ((rc = mark(k1)) != DH_OK) return rc;
        if (k1 != k23) {
            DH_CUDA(cudaEventRecord(k1_done, k1));
            DH_CUDA(cudaStreamWaitEvent(k23, k1_done, 0));
        }
        if ((rc = mark(k23)) != DH_OK) return rc;
        rc = dh_demod_process(h->demod, h->d_filt, h->filt_pitch, n, h->d_sym, h->sym_pitch, h->d_nsym, k23);
    } else {
        if ((rc = mark(k1)) != DH_OK) return rc;
        if ((rc = mark(k1)) != DH_OK) return rc;
        if ((rc = mark(k23)) != DH_OK) return rc;
        rc = dh_demod_process(h->demod, static_cast<const float*>(d_in), in_pitch, n, h->d_sym, h->sym_pitch, h->d_nsym,
                              k23);
    }
    if (rc != DH_OK) return rc;
    h->launches += (uint64_t) dh_demod_kernels_per_call(h->demod);
    if ((rc = mark(k23)) != DH_OK) return rc;
    if (k2_done) DH_CUDA(cudaEventRecord(k2_done, k23));
    if ((rc = mark(k23)) != DH_OK) return rc;
    rc = dh_decoder_process(h->decoder, h->d_sym, h->sym_pitch, h->d_nsym, h->max_syms, k23);
    if (rc != DH_OK) return rc;
    h->launches++;
    return mark(k23);
}

}  // namespace

}  // extern "C"

namespace {

int pipe_process_device(dh_pipe* h, const void* d_in_v, size_t in_pitch, size_t n, void* stream, bool s16) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_process_device: handle is NULL");
    DH_REQUIRE(n <= h->max_chunk, DH_E_INVALID, "dh_pipe_process_device: n=%zu exceeds max_chunk=%zu", n, h->max_chunk);
    DH_REQUIRE(!s16 || h->rrc, DH_E_UNSUPPORTED,
               "dh_pipe_process_device_s16: this pipe has no RRC stage to fuse the int16 conversion into");
    if (n == 0) return DH_OK;
    cudaStream_t user = (cudaStream_t) stream;
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "dh_pipe_process_device: cannot switch to device %d", h->device);
    const char* d_in = static_cast<const char*>(d_in_v);
    const size_t elem = s16 ? sizeof(int16_t) : sizeof(float);
    if (h->async_mode && h->rrc) {
        // K1(i) may start as soon as the input is ready and K2(i - 2), the previous reader of the demodulator rows
        // it alternates into, is done; K2(i) + decoder(i) follow on stream b
        const uint64_t i = h->async_step++;
        for (int k = 0; k < 4 && !h->ev_a1[3]; k++) {
            DH_CUDA(cudaEventCreateWithFlags(&h->ev_a1[k], cudaEventDisableTiming));
            DH_CUDA(cudaEventCreateWithFlags(&h->ev_a2[k], cudaEventDisableTiming));
        }
        DH_CUDA(cudaEventRecord(h->ev_start, user));
        DH_CUDA(cudaStreamWaitEvent(h->sa, h->ev_start, 0));
        if (i >= 2) DH_CUDA(cudaStreamWaitEvent(h->sa, h->ev_a2[(i - 2) & 3], 0));
        int rc = run_stages(h, d_in, in_pitch, n, h->sa, h->sb, h->ev_a1[i & 3], h->ev_a2[i & 3], s16);
        if (rc != DH_OK) return rc;
        DH_CUDA(cudaEventRecord(h->ev_done, h->sb));
        h->async_pending = true;
        h->ev_last_k1 = h->ev_a1[i & 3];
        return DH_OK;
    }
    if (!h->rrc || h->sub_chunk == 0 || n <= h->sub_chunk)
        return run_stages(h, d_in, in_pitch, n, user, user, nullptr, nullptr, s16);

    // pipelined: K1(c) on stream a; K2(c) + decoder(c) on stream b after K1(c); K1(c) may only overwrite the
    // demodulator rows it alternates into once K2(c - 2), their previous reader, is done
    const size_t sub = h->sub_chunk;
    const size_t nsub = (n + sub - 1) / sub;
    while (h->ev_k1.size() < nsub) {
        cudaEvent_t a, b;
        DH_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        DH_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_k1.push_back(a);
        h->ev_k2.push_back(b);
    }
    DH_CUDA(cudaEventRecord(h->ev_start, user));
    DH_CUDA(cudaStreamWaitEvent(h->sa, h->ev_start, 0));
    DH_CUDA(cudaStreamWaitEvent(h->sb, h->ev_start, 0));
    for (size_t c = 0; c < nsub; c++) {
        const size_t off = c * sub;
        const size_t len = n - off < sub ? n - off : sub;
        if (c >= 2) DH_CUDA(cudaStreamWaitEvent(h->sa, h->ev_k2[c - 2], 0));
        int rc = run_stages(h, d_in + off * elem, in_pitch, len, h->sa, h->sb, h->ev_k1[c], h->ev_k2[c], s16);
        if (rc != DH_OK) return rc;
    }
    // everything on stream a precedes the last K2 on stream b
    DH_CUDA(cudaEventRecord(h->ev_done, h->sb));
    DH_CUDA(cudaStreamWaitEvent(user, h->ev_done, 0));
    return DH_OK;
}

}  // namespace

extern "C" {

int dh_pipe_process_device(dh_pipe* h, const float* d_in, size_t in_pitch, size_t n, void* stream) {
    return pipe_process_device(h, d_in, in_pitch, n, stream, false);
}

int dh_pipe_process_device_s16(dh_pipe* h, const int16_t* d_in, size_t in_pitch, size_t n, void* stream) {
    return pipe_process_device(h, d_in, in_pitch, n, stream, true);
}

int dh_pipe_input_event(dh_pipe* h, void* event) {
    DH_REQUIRE(h != nullptr && event != nullptr, DH_E_INVALID, "dh_pipe_input_event: NULL argument");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    // the first kernel of a call is the only reader of the caller's block (K1, or K2 for the pipes without an RRC
    // stage); nothing else has been enqueued on its stream since
    DH_CUDA(cudaEventRecord((cudaEvent_t) event, h->last_k1_stream));
    return DH_OK;
}

int dh_pipe_set_sub_chunk(dh_pipe* h, size_t sub_chunk) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_set_sub_chunk: handle is NULL");
    DH_REQUIRE(sub_chunk % 8 == 0, DH_E_INVALID, "dh_pipe_set_sub_chunk: must be a multiple of 8");
    h->sub_chunk = sub_chunk;
    return DH_OK;
}

int dh_pipe_sync(dh_pipe* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_sync: handle is NULL");
    if (!h->async_pending) return DH_OK;
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    // stream b runs the last stage of every call and waits for stream a's K1 of the same call
    DH_CUDA(cudaStreamWaitEvent((cudaStream_t) stream, h->ev_done, 0));
    h->async_pending = false;
    return DH_OK;
}

int dh_pipe_set_async(dh_pipe* h, int enable, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_set_async: handle is NULL");
    if (!enable) {
        int rc = dh_pipe_sync(h, stream);
        if (rc != DH_OK) return rc;
    }
    h->async_mode = enable != 0;
    if (h->rrc) return dh_rrc_set_tile_preference(h->rrc, h->async_mode ? 1 : 0);
    return DH_OK;
}

int dh_pipe_discard(dh_pipe* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_discard: handle is NULL");
    // in asynchronous mode the counters belong to stream b (between the decoder kernels of consecutive calls)
    return dh_decoder_discard(h->decoder, h->async_mode && h->rrc ? (void*) h->sb : stream);
}

int dh_pipe_set_profiling(dh_pipe* h, int enable) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_set_profiling: handle is NULL");
    h->profiling = enable != 0;
    return DH_OK;
}

int dh_pipe_stage_times(dh_pipe* h, double ms[3], uint64_t* calls) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_stage_times: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    double acc[3] = {0, 0, 0};
    const size_t ncalls = h->events_used / 6;
    for (size_t c = 0; c < ncalls; c++) {
        for (int k = 0; k < 3; k++) {
            float t = 0;
            DH_CUDA(cudaEventSynchronize(h->events[6 * c + 2 * k + 1]));
            DH_CUDA(cudaEventElapsedTime(&t, h->events[6 * c + 2 * k], h->events[6 * c + 2 * k + 1]));
            acc[k] += t;
        }
    }
    h->events_used = 0;   // the events go back to the pool
    if (ms) for (int k = 0; k < 3; k++) ms[k] = acc[k];
    if (calls) *calls = ncalls;
    return DH_OK;
}

uint64_t dh_pipe_launch_count(const dh_pipe* h) { return h ? h->launches : 0; }

}  // extern "C"

namespace {

// host row pitch (elements) that allows one contiguous transfer: float32 rows are 16-byte multiples at n % 4 == 0,
// int16 rows at n % 8 == 0
size_t host_pitch_of(const dh_pipe* h, bool s16) {
    return s16 ? (h->max_chunk + 7) & ~(size_t) 7 : (h->max_chunk + 3) & ~(size_t) 3;
}

int upload(const dh_pipe* h, void* d_dst, const void* h_in, size_t in_pitch, size_t n, bool s16, cudaStream_t st) {
    const size_t elem = s16 ? sizeof(int16_t) : sizeof(float);
    const size_t pitch = host_pitch_of(h, s16);
    if (in_pitch == pitch) {
        // one contiguous transfer
        DH_CUDA(cudaMemcpyAsync(d_dst, h_in, (size_t) h->channels * in_pitch * elem, cudaMemcpyHostToDevice, st));
    } else {
        DH_CUDA(cudaMemcpy2DAsync(d_dst, pitch * elem, h_in, in_pitch * elem, n * elem, h->channels,
                                  cudaMemcpyHostToDevice, st));
    }
    return DH_OK;
}

int pipe_process_host(dh_pipe* h, const void* h_in, size_t in_pitch, size_t n, void* stream, bool s16) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_process_host: handle is NULL");
    DH_REQUIRE(n <= h->max_chunk, DH_E_INVALID, "dh_pipe_process_host: n=%zu exceeds max_chunk=%zu", n, h->max_chunk);
    if (n == 0) return DH_OK;
    DH_REQUIRE(h_in != nullptr && in_pitch >= n, DH_E_INVALID, "dh_pipe_process_host: bad input buffer");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "dh_pipe_process_host: cannot switch to device %d", h->device);
    if (!h->d_stage) {
        h->stage_pitch = (h->max_chunk + 3) & ~(size_t) 3;
        DH_CUDA(cudaMalloc(&h->d_stage, (size_t) h->channels * h->stage_pitch * sizeof(float)));
        DH_CUDA(cudaMemset(h->d_stage, 0, (size_t) h->channels * h->stage_pitch * sizeof(float)));
    }
    cudaStream_t st = (cudaStream_t) stream;
    // asynchronous mode: the staging block is still being read by K1 of the previous call (internal stream)
    if (h->async_pending && h->ev_last_k1) DH_CUDA(cudaStreamWaitEvent(st, h->ev_last_k1, 0));
    int rc = upload(h, h->d_stage, h_in, in_pitch, n, s16, st);
    if (rc != DH_OK) return rc;
    return pipe_process_device(h, h->d_stage, host_pitch_of(h, s16), n, stream, s16);
}

int pipe_submit_host(dh_pipe* h, const void* h_in, size_t in_pitch, size_t n, bool s16) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_submit_host: handle is NULL");
    DH_REQUIRE(n > 0 && n <= h->max_chunk, DH_E_INVALID, "dh_pipe_submit_host: n=%zu out of range (max_chunk=%zu)", n,
               h->max_chunk);
    DH_REQUIRE(h_in != nullptr && in_pitch >= n, DH_E_INVALID, "dh_pipe_submit_host: bad input buffer");
    DH_REQUIRE(!s16 || h->rrc, DH_E_UNSUPPORTED,
               "dh_pipe_submit_host_s16: this pipe has no RRC stage to fuse the int16 conversion into");
    DH_REQUIRE(h->submitted - h->collected < 2, DH_E_STATE,
               "dh_pipe_submit_host: two steps are already in flight, call dh_pipe_collect_step first");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "dh_pipe_submit_host: cannot switch to device %d", h->device);
    const int slot = (int) (h->submitted & 1);
    if (!h->s_copy) {
        h->stage_pitch = (h->max_chunk + 3) & ~(size_t) 3;
        DH_CUDA(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
        DH_CUDA(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
        {
            // the read-back stream runs small kernels (row compaction, counter reset) that must not queue behind the FIR
            // grid of the next step: highest priority
            int lo_prio = 0, hi_prio = 0;
            DH_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            DH_CUDA(cudaStreamCreateWithPriority(&h->s_back, cudaStreamNonBlocking, hi_prio));
        }
        for (int i = 0; i < 2; i++) {
            DH_CUDA(cudaMalloc(&h->d_slot[i], (size_t) h->channels * h->stage_pitch * sizeof(float)));
            DH_CUDA(cudaMemset(h->d_slot[i], 0, (size_t) h->channels * h->stage_pitch * sizeof(float)));
            DH_CUDA(cudaEventCreateWithFlags(&h->ev_uploaded[i], cudaEventDisableTiming));
            DH_CUDA(cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
            DH_CUDA(cudaEventCreateWithFlags(&h->ev_decoded[i], cudaEventDisableTiming));
        }
    }
    static const bool tracing = getenv("DH_PIPE_TRACE") != nullptr;
    auto trace_mark = [&](cudaStream_t st) {
        if (!tracing || h->trace.size() >= 4 * 256) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        h->trace.push_back(e);
    };
    // upload into the slot once K1 of the step that used it two submissions ago has read it
    if (h->submitted >= 2) DH_CUDA(cudaStreamWaitEvent(h->s_copy, h->ev_consumed[slot], 0));
    trace_mark(h->s_copy);
    int rc = upload(h, h->d_slot[slot], h_in, in_pitch, n, s16, h->s_copy);
    if (rc != DH_OK) return rc;
    trace_mark(h->s_copy);
    DH_CUDA(cudaEventRecord(h->ev_uploaded[slot], h->s_copy));
    DH_CUDA(cudaStreamWaitEvent(h->s_compute, h->ev_uploaded[slot], 0));
    trace_mark(h->s_compute);
    // results of this step go to the result set of its parity
    rc = dh_decoder_select_results(h->decoder, slot);
    if (rc != DH_OK) return rc;
    rc = run_stages(h, h->d_slot[slot], host_pitch_of(h, s16), n, h->s_compute, h->s_compute, nullptr, nullptr, s16);
    if (rc != DH_OK) return rc;
    trace_mark(h->s_compute);
    // K1 is the only reader of the slot, but the stages are serialised on one stream anyway
    DH_CUDA(cudaEventRecord(h->ev_consumed[slot], h->s_compute));
    DH_CUDA(cudaEventRecord(h->ev_decoded[slot], h->s_compute));
    h->submitted++;
    return DH_OK;
}

}  // namespace

extern "C" {

int dh_pipe_process_host(dh_pipe* h, const float* h_in, size_t in_pitch, size_t n, void* stream) {
    return pipe_process_host(h, h_in, in_pitch, n, stream, false);
}

int dh_pipe_process_host_s16(dh_pipe* h, const int16_t* h_in, size_t in_pitch, size_t n, void* stream) {
    return pipe_process_host(h, h_in, in_pitch, n, stream, true);
}

int dh_pipe_submit_host(dh_pipe* h, const float* h_in, size_t in_pitch, size_t n) {
    return pipe_submit_host(h, h_in, in_pitch, n, false);
}

int dh_pipe_submit_host_s16(dh_pipe* h, const int16_t* h_in, size_t in_pitch, size_t n) {
    return pipe_submit_host(h, h_in, in_pitch, n, true);
}

int dh_pipe_collect_step(dh_pipe* h) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_collect_step: handle is NULL");
    DH_REQUIRE(h->collected < h->submitted, DH_E_STATE, "dh_pipe_collect_step: nothing in flight");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    const int slot = (int) (h->collected & 1);
    DH_CUDA(cudaEventSynchronize(h->ev_decoded[slot]));
    // read-back on the copy-independent compute-side helper: the default stream would serialise with everything
    int rc = dh_decoder_collect_results(h->decoder, slot, h->s_copy_back());
    if (rc != DH_OK) return rc;
    h->collected++;
    return DH_OK;
}

int dh_pipe_collect(dh_pipe* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_collect: handle is NULL");
    int rc = dh_pipe_sync(h, stream);
    if (rc != DH_OK) return rc;
    return dh_decoder_collect(h->decoder, stream);
}

dh_decoder* dh_pipe_decoder(dh_pipe* h) { return h ? h->decoder : nullptr; }
dh_demod* dh_pipe_demod(dh_pipe* h) { return h ? h->demod : nullptr; }

// ---- state of the whole pipe: the blobs of its stages back to back behind a pipe header ----------------------------
int dh_pipe_state_size(const dh_pipe* h, size_t* bytes) {
    DH_REQUIRE(h != nullptr && bytes != nullptr, DH_E_INVALID, "dh_pipe_state_size: NULL argument");
    size_t total = sizeof(dh::StateHeader), part = 0;
    int rc = DH_OK;
    if (h->rrc) {
        rc = dh_rrc_state_size(h->rrc, &part);
        if (rc != DH_OK) return rc;
        total += part;
    }
    rc = dh_demod_state_size(h->demod, &part);
    if (rc != DH_OK) return rc;
    total += part;
    rc = dh_decoder_state_size(h->decoder, &part);
    if (rc != DH_OK) return rc;
    *bytes = total + part;
    return DH_OK;
}

int dh_pipe_state_export(dh_pipe* h, void* h_buf, size_t cap, size_t* written, void* stream) {
    DH_REQUIRE(h != nullptr && h_buf != nullptr, DH_E_INVALID, "dh_pipe_state_export: NULL argument");
    DH_REQUIRE(h->submitted == h->collected, DH_E_STATE, "dh_pipe_state_export: steps are in flight, collect them first");
    int rc = dh_pipe_sync(h, stream);   // asynchronous mode: the stages run on internal streams
    if (rc != DH_OK) return rc;
    DH_REQUIRE(cap >= sizeof(dh::StateHeader), DH_E_INVALID, "dh_pipe_state_export: buffer too small");
    char* out = static_cast<char*>(h_buf) + sizeof(dh::StateHeader);
    size_t left = cap - sizeof(dh::StateHeader), n = 0;
    if (h->rrc) {
        rc = dh_rrc_state_export(h->rrc, out, left, &n, stream);
        if (rc != DH_OK) return rc;
        out += n;
        left -= n;
    }
    rc = dh_demod_state_export(h->demod, out, left, &n, stream);
    if (rc != DH_OK) return rc;
    out += n;
    left -= n;
    rc = dh_decoder_state_export(h->decoder, out, left, &n, stream);
    if (rc != DH_OK) return rc;
    out += n;
    const size_t total = (size_t) (out - static_cast<char*>(h_buf));
    const dh::StateHeader hd = dh::make_state_header(4, h->channels, (uint32_t) h->proto, 0, 0, 0, total - sizeof(hd));
    std::memcpy(h_buf, &hd, sizeof(hd));
    if (written) *written = total;
    return DH_OK;
}

int dh_pipe_state_import(dh_pipe* h, const void* h_buf, size_t bytes, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_state_import: handle is NULL");
    DH_REQUIRE(h->submitted == h->collected, DH_E_STATE, "dh_pipe_state_import: steps are in flight, collect them first");
    const dh::StateHeader want = dh::make_state_header(4, h->channels, (uint32_t) h->proto, 0, 0, 0, 0);
    int rc = dh::check_state_header(h_buf, bytes, want, "dh_pipe_state_import");
    if (rc != DH_OK) return rc;
    rc = dh_pipe_sync(h, stream);
    if (rc != DH_OK) return rc;
    const char* in = static_cast<const char*>(h_buf) + sizeof(dh::StateHeader);
    size_t left = bytes - sizeof(dh::StateHeader);
    auto part_size = [&](size_t* n) -> int {
        DH_REQUIRE(left >= sizeof(dh::StateHeader), DH_E_INVALID, "dh_pipe_state_import: truncated blob");
        dh::StateHeader hd;
        std::memcpy(&hd, in, sizeof(hd));
        *n = sizeof(hd) + hd.payload;
        DH_REQUIRE(left >= *n, DH_E_INVALID, "dh_pipe_state_import: truncated blob");
        return DH_OK;
    };
    size_t n = 0;
    if (h->rrc) {
        if ((rc = part_size(&n)) != DH_OK) return rc;
        if ((rc = dh_rrc_state_import(h->rrc, in, n, stream)) != DH_OK) return rc;
        in += n;
        left -= n;
    }
    if ((rc = part_size(&n)) != DH_OK) return rc;
    if ((rc = dh_demod_state_import(h->demod, in, n, stream)) != DH_OK) return rc;
    in += n;
    left -= n;
    if ((rc = part_size(&n)) != DH_OK) return rc;
    return dh_decoder_state_import(h->decoder, in, n, stream);
}

int dh_pipe_last_symbols(dh_pipe* h, const uint8_t** d_sym, size_t* sym_pitch, const uint32_t** d_nsym) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_last_symbols: handle is NULL");
    if (d_sym) *d_sym = h->d_sym;
    if (sym_pitch) *sym_pitch = h->sym_pitch;
    if (d_nsym) *d_nsym = h->d_nsym;
    return DH_OK;
}

int dh_pipe_read_symbols(dh_pipe* h, uint32_t channel, uint8_t* h_buf, size_t cap, size_t* count) {
    DH_REQUIRE(h != nullptr && channel < h->channels, DH_E_INVALID, "dh_pipe_read_symbols: bad handle or channel");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    if (h->async_pending) DH_CUDA(cudaStreamSynchronize(h->sb));
    uint32_t n = 0;
    DH_CUDA(cudaMemcpy(&n, h->d_nsym + channel, sizeof(n), cudaMemcpyDeviceToHost));
    if (count) *count = n;
    const size_t take = n < cap ? n : cap;
    if (take && h_buf) {
        DH_CUDA(cudaMemcpy(h_buf, h->d_sym + (size_t) channel * h->sym_pitch, take, cudaMemcpyDeviceToHost));
    }
    return DH_OK;
}

size_t dh_pipe_host_pitch(const dh_pipe* h) { return h ? host_pitch_of(h, false) : 0; }

size_t dh_pipe_host_pitch_s16(const dh_pipe* h) { return h ? host_pitch_of(h, true) : 0; }

uint32_t dh_pipe_channels(const dh_pipe* h) { return h ? h->channels : 0; }

void dh_pipe_destroy(dh_pipe* h) {
    if (!h) return;
    {
        dh::DeviceGuard guard(h->device);
        if (!h->trace.empty()) {
            cudaDeviceSynchronize();
            const size_t steps = h->trace.size() / 4;
            for (size_t k = 0; k < steps; k++) {
                float t[4] = {0, 0, 0, 0};
                for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], h->trace[0], h->trace[4 * k + j]);
                fprintf(stderr, "[pipe trace] step %3zu: upload %9.3f .. %9.3f ms, kernels %9.3f .. %9.3f ms\n", k, t[0], t[1], t[2], t[3]);
            }
            for (cudaEvent_t e : h->trace) cudaEventDestroy(e);
        }
        for (cudaEvent_t e : h->events) cudaEventDestroy(e);
        for (cudaEvent_t e : h->ev_k1) cudaEventDestroy(e);
        for (cudaEvent_t e : h->ev_k2) cudaEventDestroy(e);
        for (int k = 0; k < 4; k++) {
            if (h->ev_a1[k]) cudaEventDestroy(h->ev_a1[k]);
            if (h->ev_a2[k]) cudaEventDestroy(h->ev_a2[k]);
        }
        if (h->ev_start) cudaEventDestroy(h->ev_start);
        if (h->ev_done) cudaEventDestroy(h->ev_done);
        if (h->sa) cudaStreamDestroy(h->sa);
        if (h->sb) cudaStreamDestroy(h->sb);
        if (h->s_copy) cudaStreamDestroy(h->s_copy);
        if (h->s_compute) cudaStreamDestroy(h->s_compute);
        if (h->s_back) cudaStreamDestroy(h->s_back);
        for (int i = 0; i < 2; i++) {
            if (h->ev_uploaded[i]) cudaEventDestroy(h->ev_uploaded[i]);
            if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
            if (h->ev_decoded[i]) cudaEventDestroy(h->ev_decoded[i]);
            cudaFree(h->d_slot[i]);
        }
    }
    dh_rrc_destroy(h->rrc);
    dh_demod_destroy(h->demod);
    dh_decoder_destroy(h->decoder);
    {
        dh::DeviceGuard guard(h->device);
        cudaFree(h->d_nsym);
        cudaFree(h->d_stage);
    }
    delete h;
}

}  // extern "C"
