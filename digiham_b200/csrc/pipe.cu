// pipe.cu — dh_pipe_*: the whole per-channel pipe of the reference's example scripts as one bank object.
//
//   DMR / YSF : WideRrcFilter -> GfskDemodulator(10) -> decoder   (reference examples/dmr-decoder.sh:19-23,
//                                                                  examples/ysf-decoder.sh:19-23)
//   POCSAG    : FskDemodulator(40, invert) -> decoder             (reference examples/pocsag-decoder.sh:19-21)
//
// In the reference every `|` is a process boundary with a 1024-item ring in between (src/lib/cli.cpp:10,101-106);
// here the stages are kernels on one stream and each stage writes straight into the next stage's carry-aware
// input rows (dh_demod_reserve / dh_decoder_reserve), so no intermediate copy exists.
#include "common.cuh"

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
    float* d_stage = nullptr;     // device staging for host input
    size_t stage_pitch = 0;
    // optional per-stage device timing (CUDA events on the caller's stream)
    bool profiling = false;
    std::vector<cudaEvent_t> events;   // 4 per process call: start, after K1, after K2, after decoder
    uint64_t launches = 0;             // kernels launched by this pipe since creation
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
    if (!pocsag) rc = dh_rrc_create(&h->rrc, device, channels, DH_RRC_WIDE);
    if (rc == DH_OK) rc = pocsag ? dh_demod_create(&h->demod, device, channels, 0, 40, 1)
                                 : dh_demod_create(&h->demod, device, channels, 1, 10, 0);
    if (rc == DH_OK) rc = dh_decoder_create(&h->decoder, device, channels, proto);
    if (rc == DH_OK) rc = dh_demod_reserve(h->demod, max_chunk, &h->d_filt, &h->filt_pitch);
    if (rc == DH_OK) {
        h->max_syms = dh_demod_max_symbols(h->demod, max_chunk);
        rc = dh_decoder_reserve(h->decoder, h->max_syms, &h->d_sym, &h->sym_pitch);
    }
    if (rc == DH_OK) {
        dh::DeviceGuard guard(device);
        cudaError_t e = cudaMalloc(&h->d_nsym, (size_t) channels * sizeof(uint32_t));
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

int dh_pipe_process_device(dh_pipe* h, const float* d_in, size_t in_pitch, size_t n, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_process_device: handle is NULL");
    DH_REQUIRE(n <= h->max_chunk, DH_E_INVALID, "dh_pipe_process_device: n=%zu exceeds max_chunk=%zu", n, h->max_chunk);
    if (n == 0) return DH_OK;
    cudaStream_t st = (cudaStream_t) stream;
    auto mark = [&]() -> int {
        if (!h->profiling) return DH_OK;
        dh::DeviceGuard guard(h->device);
        cudaEvent_t e;
        DH_CUDA(cudaEventCreate(&e));
        DH_CUDA(cudaEventRecord(e, st));
        h->events.push_back(e);
        return DH_OK;
    };
    int rc = mark();
    if (rc != DH_OK) return rc;
    if (h->rrc) {
        // the demodulator's input rows alternate between two buffers from call to call
        rc = dh_demod_reserve(h->demod, h->max_chunk, &h->d_filt, &h->filt_pitch);
        if (rc != DH_OK) return rc;
        rc = dh_rrc_process(h->rrc, d_in, in_pitch, h->d_filt, h->filt_pitch, n, stream);
        if (rc != DH_OK) return rc;
        h->launches++;
        if ((rc = mark()) != DH_OK) return rc;
        rc = dh_demod_process(h->demod, h->d_filt, h->filt_pitch, n, h->d_sym, h->sym_pitch, h->d_nsym, stream);
    } else {
        if ((rc = mark()) != DH_OK) return rc;
        rc = dh_demod_process(h->demod, d_in, in_pitch, n, h->d_sym, h->sym_pitch, h->d_nsym, stream);
    }
    if (rc != DH_OK) return rc;
    h->launches++;
    if ((rc = mark()) != DH_OK) return rc;
    rc = dh_decoder_process(h->decoder, h->d_sym, h->sym_pitch, h->d_nsym, h->max_syms, stream);
    if (rc != DH_OK) return rc;
    h->launches++;
    return mark();
}

int dh_pipe_set_profiling(dh_pipe* h, int enable) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_set_profiling: handle is NULL");
    h->profiling = enable != 0;
    return DH_OK;
}

int dh_pipe_stage_times(dh_pipe* h, double ms[3], uint64_t* calls) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_stage_times: handle is NULL");
    dh::DeviceGuard guard(h->device);
    double acc[3] = {0, 0, 0};
    const size_t ncalls = h->events.size() / 4;
    for (size_t c = 0; c < ncalls; c++) {
        DH_CUDA(cudaEventSynchronize(h->events[4 * c + 3]));
        for (int k = 0; k < 3; k++) {
            float t = 0;
            DH_CUDA(cudaEventElapsedTime(&t, h->events[4 * c + k], h->events[4 * c + k + 1]));
            acc[k] += t;
        }
    }
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    h->events.clear();
    if (ms) for (int k = 0; k < 3; k++) ms[k] = acc[k];
    if (calls) *calls = ncalls;
    return DH_OK;
}

uint64_t dh_pipe_launch_count(const dh_pipe* h) { return h ? h->launches : 0; }

int dh_pipe_process_host(dh_pipe* h, const float* h_in, size_t in_pitch, size_t n, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_process_host: handle is NULL");
    DH_REQUIRE(n <= h->max_chunk, DH_E_INVALID, "dh_pipe_process_host: n=%zu exceeds max_chunk=%zu", n, h->max_chunk);
    if (n == 0) return DH_OK;
    DH_REQUIRE(h_in != nullptr && in_pitch >= n, DH_E_INVALID, "dh_pipe_process_host: bad input buffer");
    dh::DeviceGuard guard(h->device);
    if (!h->d_stage) {
        h->stage_pitch = (h->max_chunk + 3) & ~(size_t) 3;
        DH_CUDA(cudaMalloc(&h->d_stage, (size_t) h->channels * h->stage_pitch * sizeof(float)));
        DH_CUDA(cudaMemset(h->d_stage, 0, (size_t) h->channels * h->stage_pitch * sizeof(float)));
    }
    cudaStream_t st = (cudaStream_t) stream;
    if (in_pitch == h->stage_pitch) {
        // one contiguous transfer
        DH_CUDA(cudaMemcpyAsync(h->d_stage, h_in, (size_t) h->channels * in_pitch * sizeof(float),
                                cudaMemcpyHostToDevice, st));
    } else {
        DH_CUDA(cudaMemcpy2DAsync(h->d_stage, h->stage_pitch * sizeof(float), h_in, in_pitch * sizeof(float),
                                  n * sizeof(float), h->channels, cudaMemcpyHostToDevice, st));
    }
    return dh_pipe_process_device(h, h->d_stage, h->stage_pitch, n, stream);
}

int dh_pipe_collect(dh_pipe* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_pipe_collect: handle is NULL");
    return dh_decoder_collect(h->decoder, stream);
}

dh_decoder* dh_pipe_decoder(dh_pipe* h) { return h ? h->decoder : nullptr; }

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
    uint32_t n = 0;
    DH_CUDA(cudaMemcpy(&n, h->d_nsym + channel, sizeof(n), cudaMemcpyDeviceToHost));
    if (count) *count = n;
    const size_t take = n < cap ? n : cap;
    if (take && h_buf) {
        DH_CUDA(cudaMemcpy(h_buf, h->d_sym + (size_t) channel * h->sym_pitch, take, cudaMemcpyDeviceToHost));
    }
    return DH_OK;
}

size_t dh_pipe_host_pitch(const dh_pipe* h) { return h ? (h->max_chunk + 3) & ~(size_t) 3 : 0; }

void dh_pipe_destroy(dh_pipe* h) {
    if (!h) return;
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
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
