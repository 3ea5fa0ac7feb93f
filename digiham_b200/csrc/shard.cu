// shard.cu — dh_shard_*: one protocol pipe sharded over the ranks of an NCCL communicator (one process per GPU).
//
// Channels never interact (one module instance per channel in the reference, examples/dmr-decoder.sh:19-23), so
// rank r of R owns the contiguous channel range [r*N/R, (r+1)*N/R) and runs an ordinary dh_pipe on it.  The only
// exchange steps are the two the ingest layout forces (BASELINE configs[3]):
//   scatter  the ingest (root) rank holds the sample block of ALL channels and hands every peer its rows (float32 or
//            int16 samples): the root maps the peers' input slots through CUDA IPC and writes the rows with one
//            copy-engine transfer per peer over NVLink (no SM is spent on it, all peers' copies run side by side),
//            NCCL only carries the two 4-byte tokens per peer and step that order them ("slot free" / "rows landed");
//            grouped ncclSend / ncclRecv of the rows themselves is the fallback when the slots cannot be mapped;
//   gather   every rank packs the decoder results of the step — fixed-slot byte rows, 16-byte metadata event records,
//            per-channel counts — into one wire block with per-step slot widths; on a peer the pack kernel stores the
//            block straight into the root's wire buffer (mapped through CUDA IPC: pack and gather are one kernel, only
//            the used bytes cross NVLink), ordered by two 4-byte NCCL tokens; ncclSend / ncclRecv of the block is the
//            fallback.  The root feeds all blocks through one host-side result sink (result_sink.cu) in global
//            channel order.
// The three phases of consecutive steps overlap: scatter(k+1) runs on its own stream and communicator while the
// kernels of step k run (cross-step pipelined, dh_pipe_set_async) and the wire block of step k-1 travels on a third
// stream over a second communicator (ncclCommSplit), so neither collective waits behind the other.  Input blocks,
// result sets and wire blocks are double-buffered; CUDA events carry exactly the reuse dependencies.
//
// NCCL is resolved at run time from the library already loaded in the process (the one that created the
// communicator the host passes in), falling back to libnccl.so.2 on the loader path.
#include "decoder_ops.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    char why[200] = "";
};

const NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            snprintf(api.why, sizeof(api.why), "cannot load libnccl.so.2: %s", dlerror());
            return;
        }
        bool all = true;
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(lib, name);
            if (!p) {
                all = false;
                snprintf(api.why, sizeof(api.why), "libnccl lacks %s (NCCL >= 2.18 is required)", name);
            }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
        api.CommSplit = (decltype(api.CommSplit)) sym("ncclCommSplit");
        api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
        api.CommCount = (decltype(api.CommCount)) sym("ncclCommCount");
        api.CommUserRank = (decltype(api.CommUserRank)) sym("ncclCommUserRank");
        api.GroupStart = (decltype(api.GroupStart)) sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
        api.Send = (decltype(api.Send)) sym("ncclSend");
        api.Recv = (decltype(api.Recv)) sym("ncclRecv");
        api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
        api.ok = all;
    });
    return api;
}

#define DH_NCCL(call)                                                                                        \
    do {                                                                                                     \
        ncclResult_t r__ = (call);                                                                           \
        if (r__ != ncclSuccess) {                                                                            \
            dh::set_error("%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r__));   \
            return DH_E_NCCL;                                                                                \
        }                                                                                                    \
    } while (0)

#define DH_NEED_NCCL()                                                            \
    do {                                                                          \
        if (!nccl().ok) {                                                         \
            dh::set_error("NCCL is not available: %s", nccl().why);               \
            return DH_E_UNSUPPORTED;                                              \
        }                                                                         \
    } while (0)

inline size_t round16(size_t v) { return (v + 15) & ~(size_t) 15; }

// contiguous, balanced split: the first (total % world) ranks own one extra channel
void range_of(uint64_t total, int world, int rank, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = total / (uint64_t) world, extra = total % (uint64_t) world;
    const uint64_t r = (uint64_t) rank;
    *lo = r * base + (r < extra ? r : extra);
    *hi = *lo + base + (r < extra ? 1 : 0);
}

// Wire block of n channels: counts [3][n] u32 (out_len, ev_len, flags) | byte rows [n][w_out] | event rows [n][w_ev]
struct WireLayout {
    uint32_t w_out = 0;   // bytes per channel and step, multiple of 16
    uint32_t w_ev = 0;    // event records per channel and step
    size_t off_out(uint32_t n) const { return round16(3 * (size_t) n * sizeof(uint32_t)); }
    size_t off_ev(uint32_t n) const { return off_out(n) + (size_t) n * w_out; }
    size_t bytes(uint32_t n) const { return off_ev(n) + (size_t) n * w_ev * sizeof(dh::DecEvent); }
};

// One warp per channel: copies the used part of the channel's result slots into the wire block in 16-byte units,
// writes the (clamped) counts into the wire header and resets the bank's counters for the next step that uses the set.
__global__ void __launch_bounds__(128) pack_results_kernel(uint32_t* __restrict__ counts, const uint8_t* __restrict__ out,
                                                           uint32_t out_cap, const dh::DecEvent* __restrict__ ev,
                                                           uint32_t ev_cap, uint8_t* __restrict__ wire, uint32_t n,
                                                           uint32_t w_out, uint32_t w_ev, size_t off_out, size_t off_ev) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t* wc = reinterpret_cast<uint32_t*>(wire);
    for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += warps) {
        uint32_t out_len = counts[c], ev_len = counts[n + c], flags = counts[2 * (size_t) n + c];
        if (out_len > w_out) {
            out_len = w_out;
            flags |= dh::kFlagOutOverflow;
        }
        if (ev_len > w_ev) {
            ev_len = w_ev;
            flags |= dh::kFlagEventOverflow;
        }
        __syncwarp();
        if (lane == 0) {
            wc[c] = out_len;
            wc[n + c] = ev_len;
            wc[2 * (size_t) n + c] = flags;
            counts[c] = 0;
            counts[n + c] = 0;
            counts[2 * (size_t) n + c] = 0;
        }
        const uint4* so = reinterpret_cast<const uint4*>(out + (size_t) c * out_cap);
        uint4* dst_o = reinterpret_cast<uint4*>(wire + off_out + (size_t) c * w_out);
        for (uint32_t i = lane; i < (out_len + 15) / 16; i += 32) dst_o[i] = so[i];
        const uint4* se = reinterpret_cast<const uint4*>(ev + (size_t) c * ev_cap);
        uint4* dst_e = reinterpret_cast<uint4*>(wire + off_ev + (size_t) c * w_ev * sizeof(dh::DecEvent));
        for (uint32_t i = lane; i < ev_len; i += 32) dst_e[i] = se[i];
    }
}

}  // namespace

struct dh_shard {
    int device = 0, rank = 0, world = 1, root = 0, proto = 0, fmt = 0;
    uint64_t channels_total = 0;
    uint32_t n_local = 0;
    size_t max_chunk = 0, pitch = 0, elem = 4;
    ncclComm_t comm_in = nullptr, comm_out = nullptr;
    dh_pipe* pipe = nullptr;
    dh_decoder* dec = nullptr;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr, s_back = nullptr;
    void* d_slot[2] = {nullptr, nullptr};          // received input blocks (ranks other than the root)
    // scatter through peer memory: the root's views of every peer's two slots (cudaIpcOpenMemHandle), one copy stream
    // per peer so that the copy engines work on all peers at once, and the 4-byte NCCL tokens that order the copies
    bool ipc_scatter = false;
    // gather through peer memory: the peers' views of the root's two wire buffers; their pack kernel stores its wire
    // block straight into the root's memory over NVLink (pack + gather in one kernel), NCCL carries two tokens
    bool ipc_gather = false;
    void* root_wire[2] = {nullptr, nullptr};
    std::vector<void*> peer_slot[2];
    std::vector<cudaStream_t> s_peer;
    std::vector<cudaEvent_t> ev_peer;
    cudaEvent_t ev_ready = nullptr;
    uint32_t* d_token = nullptr;                    // [4 * world] scratch for the tokens (scatter: first half, gather: second)
    std::vector<cudaEvent_t> trace;                 // diagnostics (DH_SHARD_TRACE): 6 timing events per step
    cudaEvent_t ev_user = nullptr, ev_scattered = nullptr, ev_computed = nullptr, ev_join = nullptr;
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};   // first kernel of the step has read its input block
    cudaEvent_t ev_packed[2] = {nullptr, nullptr};     // result set of the step has been packed (and reset)
    cudaEvent_t ev_gathered[2] = {nullptr, nullptr};   // wire block of the step has left / arrived
    WireLayout wire;
    uint8_t* d_wire[2] = {nullptr, nullptr};       // root: the regions of all ranks back to back; others: own block
    std::vector<uint64_t> lo_of;
    std::vector<uint32_t> n_of;
    std::vector<size_t> region_off;
    dh::ResultSink sink;                            // root only: all channels
    bool sink_ready = false;
    uint64_t submitted = 0, collected = 0, packs = 0;
};

namespace {

// Collective (all ranks, during dh_shard_create): the peers publish CUDA IPC handles of their two input slots, the root
// maps them and tells everybody whether the scatter can go through peer memory.  Any failure on the way (no peer
// access, IPC refused by the platform, DH_SHARD_NO_IPC set) just selects the NCCL data path.
int setup_ipc_scatter(dh_shard* h) {
    const int world = h->world, rank = h->rank, root = h->root;
    const bool is_root = rank == root;
    DH_CUDA(cudaMalloc(&h->d_token, 4 * (size_t) world * sizeof(uint32_t)));
    DH_CUDA(cudaMemset(h->d_token, 0, 4 * (size_t) world * sizeof(uint32_t)));
    DH_CUDA(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    struct Handles {
        cudaIpcMemHandle_t slot[2];
        uint32_t ok;
        uint32_t pad[3];
    };
    Handles* d_handles = nullptr;
    DH_CUDA(cudaMalloc(&d_handles, (size_t) world * sizeof(Handles)));
    std::vector<Handles> all((size_t) world);
    uint32_t usable = getenv("DH_SHARD_NO_IPC") ? 0u : 1u;
    if (!is_root) {
        Handles mine;
        std::memset(&mine, 0, sizeof(mine));
        mine.ok = usable;
        for (int i = 0; i < 2 && mine.ok; i++) {
            if (cudaIpcGetMemHandle(&mine.slot[i], h->d_slot[i]) != cudaSuccess) {
                cudaGetLastError();
                mine.ok = 0;
            }
        }
        DH_CUDA(cudaMemcpyAsync(d_handles + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->s_in));
        DH_NCCL(nccl().Send(d_handles + rank, sizeof(Handles), ncclInt8, root, h->comm_in, h->s_in));
        // the root's verdict on the scatter + the handles of its two wire buffers (ok = may be used for the gather)
        Handles theirs;
        DH_NCCL(nccl().Recv(d_handles + root, sizeof(Handles), ncclInt8, root, h->comm_in, h->s_in));
        DH_CUDA(cudaMemcpyAsync(&theirs, d_handles + root, sizeof(Handles), cudaMemcpyDeviceToHost, h->s_in));
        DH_CUDA(cudaStreamSynchronize(h->s_in));
        usable = theirs.pad[0];
        uint32_t gather_ok = theirs.ok;
        for (int i = 0; i < 2 && gather_ok; i++) {
            if (cudaIpcOpenMemHandle(&h->root_wire[i], theirs.slot[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                h->root_wire[i] = nullptr;
                gather_ok = 0;
            }
        }
        // every peer reports, the root answers with the common decision
        DH_CUDA(cudaMemcpyAsync(h->d_token + 1, &gather_ok, sizeof(uint32_t), cudaMemcpyHostToDevice, h->s_in));
        DH_NCCL(nccl().Send(h->d_token + 1, sizeof(uint32_t), ncclInt8, root, h->comm_in, h->s_in));
        DH_NCCL(nccl().Recv(h->d_token + 1, sizeof(uint32_t), ncclInt8, root, h->comm_in, h->s_in));
        DH_CUDA(cudaMemcpyAsync(&gather_ok, h->d_token + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->s_in));
        DH_CUDA(cudaStreamSynchronize(h->s_in));
        if (!gather_ok) {
            for (int i = 0; i < 2; i++) {
                if (h->root_wire[i]) cudaIpcCloseMemHandle(h->root_wire[i]);
                h->root_wire[i] = nullptr;
            }
        }
        h->ipc_gather = gather_ok != 0;
    } else {
        DH_NCCL(nccl().GroupStart());
        for (int r = 0; r < world; r++)
            if (r != root) DH_NCCL(nccl().Recv(d_handles + r, sizeof(Handles), ncclInt8, r, h->comm_in, h->s_in));
        DH_NCCL(nccl().GroupEnd());
        DH_CUDA(cudaMemcpyAsync(all.data(), d_handles, (size_t) world * sizeof(Handles), cudaMemcpyDeviceToHost, h->s_in));
        DH_CUDA(cudaStreamSynchronize(h->s_in));
        for (int i = 0; i < 2; i++) h->peer_slot[i].assign((size_t) world, nullptr);
        for (int r = 0; r < world && usable; r++) {
            if (r == root) continue;
            if (!all[r].ok) usable = 0;
            for (int i = 0; i < 2 && usable; i++) {
                if (cudaIpcOpenMemHandle(&h->peer_slot[i][r], all[r].slot[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    h->peer_slot[i][r] = nullptr;
                    usable = 0;
                }
            }
        }
        if (usable) {
            h->s_peer.assign((size_t) world, nullptr);
            h->ev_peer.assign((size_t) world, nullptr);
            for (int r = 0; r < world; r++) {
                if (r == root) continue;
                DH_CUDA(cudaStreamCreateWithFlags(&h->s_peer[r], cudaStreamNonBlocking));
                DH_CUDA(cudaEventCreateWithFlags(&h->ev_peer[r], cudaEventDisableTiming));
            }
        }
        // the root's own wire buffers for the gather direction
        Handles mine;
        std::memset(&mine, 0, sizeof(mine));
        mine.ok = (getenv("DH_SHARD_NO_IPC") || getenv("DH_SHARD_NO_IPC_GATHER")) ? 0u : 1u;
        for (int i = 0; i < 2 && mine.ok; i++) {
            if (cudaIpcGetMemHandle(&mine.slot[i], h->d_wire[i]) != cudaSuccess) {
                cudaGetLastError();
                mine.ok = 0;
            }
        }
        mine.pad[0] = usable;   // the verdict on the scatter travels in the same message
        DH_CUDA(cudaMemcpyAsync(d_handles + root, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->s_in));
        DH_NCCL(nccl().GroupStart());
        for (int r = 0; r < world; r++)
            if (r != root) DH_NCCL(nccl().Send(d_handles + root, sizeof(Handles), ncclInt8, r, h->comm_in, h->s_in));
        DH_NCCL(nccl().GroupEnd());
        // collect the peers' reports on the gather mapping, answer with the common decision
        DH_NCCL(nccl().GroupStart());
        for (int r = 0; r < world; r++)
            if (r != root) DH_NCCL(nccl().Recv(h->d_token + world + r, sizeof(uint32_t), ncclInt8, r, h->comm_in, h->s_in));
        DH_NCCL(nccl().GroupEnd());
        std::vector<uint32_t> reports((size_t) world, 1);
        DH_CUDA(cudaMemcpyAsync(reports.data(), h->d_token + world, (size_t) world * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                h->s_in));
        DH_CUDA(cudaStreamSynchronize(h->s_in));
        uint32_t gather_ok = mine.ok;
        for (int r = 0; r < world; r++)
            if (r != root && !reports[r]) gather_ok = 0;
        DH_CUDA(cudaMemcpyAsync(h->d_token + 1, &gather_ok, sizeof(uint32_t), cudaMemcpyHostToDevice, h->s_in));
        DH_NCCL(nccl().GroupStart());
        for (int r = 0; r < world; r++)
            if (r != root) DH_NCCL(nccl().Send(h->d_token + 1, sizeof(uint32_t), ncclInt8, r, h->comm_in, h->s_in));
        DH_NCCL(nccl().GroupEnd());
        DH_CUDA(cudaStreamSynchronize(h->s_in));
        h->ipc_gather = gather_ok != 0;
        DH_CUDA(cudaMemsetAsync(h->d_token, 0, 4 * (size_t) world * sizeof(uint32_t), h->s_in));
        DH_CUDA(cudaStreamSynchronize(h->s_in));
    }
    cudaFree(d_handles);
    h->ipc_scatter = usable != 0;
    return DH_OK;
}

}  // namespace

extern "C" {

void dh_shard_destroy(dh_shard* h);

int dh_shard_unique_id(uint8_t id[128]) {
    DH_REQUIRE(id != nullptr, DH_E_INVALID, "dh_shard_unique_id: id is NULL");
    DH_NEED_NCCL();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    DH_NCCL(nccl().GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return DH_OK;
}

int dh_shard_comm_init(void** comm, const uint8_t id[128], int rank, int world, int device) {
    DH_REQUIRE(comm != nullptr && id != nullptr, DH_E_INVALID, "dh_shard_comm_init: NULL argument");
    *comm = nullptr;
    DH_REQUIRE(world >= 1 && rank >= 0 && rank < world, DH_E_INVALID, "dh_shard_comm_init: bad rank %d / world %d", rank,
               world);
    DH_NEED_NCCL();
    DH_CUDA(cudaSetDevice(device));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    DH_NCCL(nccl().CommInitRank(&c, world, u, rank));
    *comm = c;
    return DH_OK;
}

int dh_shard_comm_destroy(void* comm) {
    if (!comm) return DH_OK;
    DH_NEED_NCCL();
    DH_NCCL(nccl().CommDestroy((ncclComm_t) comm));
    return DH_OK;
}

int dh_shard_channel_range(uint64_t channels_total, int world, int rank, uint64_t* lo, uint64_t* hi) {
    DH_REQUIRE(world >= 1 && rank >= 0 && rank < world && lo && hi, DH_E_INVALID, "dh_shard_channel_range: bad argument");
    range_of(channels_total, world, rank, lo, hi);
    return DH_OK;
}

int dh_shard_wire_layout(int proto, size_t max_chunk, uint32_t channels, uint32_t* slot_bytes, uint32_t* slot_events,
                         size_t* block_bytes) {
    const dh::ProtoOps* ops = dh::proto_ops(proto);
    DH_REQUIRE(ops != nullptr, DH_E_UNSUPPORTED, "dh_shard_wire_layout: protocol %d not supported", proto);
    DH_REQUIRE(max_chunk > 0, DH_E_INVALID, "dh_shard_wire_layout: max_chunk must be > 0");
    // symbols one step can append per channel: the demodulator's own bound (dh_demod_max_symbols: carried samples +
    // chunk at sps - 1 samples per symbol), rounded like the decoder bank's rows (dh_decoder_reserve)
    const size_t sps = proto == DH_PROTO_POCSAG ? 40 : (proto == DH_PROTO_NXDN ? 20 : 10);
    const size_t carry = (size_t) 100 * sps + 16;
    const size_t max_syms = ((carry + max_chunk) / (sps - 1) + 2 + 15) & ~(size_t) 15;
    WireLayout w;
    w.w_out = (uint32_t) round16(ops->out_bytes(max_syms));
    w.w_ev = ops->events(max_syms);
    if (slot_bytes) *slot_bytes = w.w_out;
    if (slot_events) *slot_events = w.w_ev;
    if (block_bytes) *block_bytes = w.bytes(channels);
    return DH_OK;
}

int dh_shard_create(dh_shard** out, void* nccl_comm, int rank, int world, int root, int device, uint64_t channels_total,
                    int proto, size_t max_chunk, int sample_format) {
    DH_REQUIRE(out != nullptr, DH_E_INVALID, "dh_shard_create: out is NULL");
    *out = nullptr;
    DH_REQUIRE(nccl_comm != nullptr || world == 1, DH_E_INVALID, "dh_shard_create: communicator is NULL");
    DH_REQUIRE(world >= 1 && rank >= 0 && rank < world && root >= 0 && root < world, DH_E_INVALID,
               "dh_shard_create: bad rank %d / world %d / root %d", rank, world, root);
    DH_REQUIRE(channels_total >= (uint64_t) world, DH_E_INVALID, "dh_shard_create: fewer channels than ranks");
    DH_REQUIRE(sample_format == DH_FMT_F32 || sample_format == DH_FMT_S16, DH_E_INVALID,
               "dh_shard_create: unknown sample format %d", sample_format);
    DH_REQUIRE(max_chunk > 0, DH_E_INVALID, "dh_shard_create: max_chunk must be > 0");
    if (world > 1) {
        DH_NEED_NCCL();
        int cnt = 0, me = -1;
        DH_NCCL(nccl().CommCount((ncclComm_t) nccl_comm, &cnt));
        DH_NCCL(nccl().CommUserRank((ncclComm_t) nccl_comm, &me));
        DH_REQUIRE(cnt == world && me == rank, DH_E_INVALID,
                   "dh_shard_create: the communicator has rank %d of %d, the call says %d of %d", me, cnt, rank, world);
    }
    dh_shard* h = new (std::nothrow) dh_shard();
    DH_REQUIRE(h != nullptr, DH_E_NOMEM, "dh_shard_create: out of host memory");
    h->device = device;
    h->rank = rank;
    h->world = world;
    h->root = root;
    h->proto = proto;
    h->fmt = sample_format;
    h->channels_total = channels_total;
    h->max_chunk = max_chunk;
    h->elem = sample_format == DH_FMT_S16 ? sizeof(int16_t) : sizeof(float);
    h->pitch = sample_format == DH_FMT_S16 ? (max_chunk + 7) & ~(size_t) 7 : (max_chunk + 3) & ~(size_t) 3;
    h->lo_of.resize(world);
    h->n_of.resize(world);
    for (int r = 0; r < world; r++) {
        uint64_t lo, hi;
        range_of(channels_total, world, r, &lo, &hi);
        h->lo_of[r] = lo;
        h->n_of[r] = (uint32_t) (hi - lo);
    }
    h->n_local = h->n_of[rank];
    auto fail = [&](int rc) {
        dh_shard_destroy(h);
        return rc;
    };
    int rc = dh_pipe_create(&h->pipe, device, h->n_local, proto, max_chunk);
    if (rc != DH_OK) return fail(rc);
    h->dec = dh_pipe_decoder(h->pipe);
    dh::DecoderView view;
    rc = dh::decoder_view(h->dec, 0, &view);
    if (rc != DH_OK) return fail(rc);
    const dh::ProtoOps* ops = dh::proto_ops(proto);
    h->wire.w_out = (uint32_t) round16(ops->out_bytes(view.max_syms));
    h->wire.w_ev = ops->events(view.max_syms);

    dh::DeviceGuard guard(device);
    if (!guard.ok) {
        dh::set_error("dh_shard_create: cannot switch to device %d", device);
        return fail(DH_E_NODEVICE);
    }
#define DH_TRY(call)                                                                          \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            dh::set_error("dh_shard_create: %s -> %s", #call, cudaGetErrorString(e__));       \
            return fail((int) e__);                                                           \
        }                                                                                     \
    } while (0)
    // The communication streams get the highest priority: a NCCL kernel launched behind an 80 000-CTA FIR grid of
    // equal priority is only dispatched in that grid's tail, which turned the 4-byte ordering tokens of the scatter
    // into 2-3 ms waits (DH_SHARD_TRACE timeline)
    int lo_prio = 0, hi_prio = 0;
    DH_TRY(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    DH_TRY(cudaStreamCreateWithPriority(&h->s_in, cudaStreamNonBlocking, hi_prio));
    DH_TRY(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
    DH_TRY(cudaStreamCreateWithPriority(&h->s_out, cudaStreamNonBlocking, hi_prio));
    DH_TRY(cudaStreamCreateWithPriority(&h->s_back, cudaStreamNonBlocking, hi_prio));   // read-back: compaction kernels + copies
    DH_TRY(cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
    DH_TRY(cudaEventCreateWithFlags(&h->ev_scattered, cudaEventDisableTiming));
    DH_TRY(cudaEventCreateWithFlags(&h->ev_computed, cudaEventDisableTiming));
    DH_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->region_off.resize(world);
    size_t total = 0;
    for (int r = 0; r < world; r++) {
        h->region_off[r] = total;
        total += round16(h->wire.bytes(h->n_of[r]));
    }
    for (int i = 0; i < 2; i++) {
        DH_TRY(cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
        DH_TRY(cudaEventCreateWithFlags(&h->ev_packed[i], cudaEventDisableTiming));
        // blocking sync: the ranks that only wait in dh_shard_collect_step sleep instead of spinning on a core the
        // gathering rank's replay threads could use
        DH_TRY(cudaEventCreateWithFlags(&h->ev_gathered[i], cudaEventDisableTiming | cudaEventBlockingSync));
        if (rank != root) {
            DH_TRY(cudaMalloc(&h->d_slot[i], (size_t) h->n_local * h->pitch * h->elem));
            DH_TRY(cudaMemset(h->d_slot[i], 0, (size_t) h->n_local * h->pitch * h->elem));
        }
        const size_t wb = rank == root ? total : h->wire.bytes(h->n_local);
        DH_TRY(cudaMalloc(&h->d_wire[i], wb));
        DH_TRY(cudaMemset(h->d_wire[i], 0, wb));
    }
#undef DH_TRY
    rc = dh_pipe_set_async(h->pipe, 1, h->s_cmp);
    if (rc != DH_OK) return fail(rc);
    if (rank == root) {
        DH_REQUIRE(channels_total <= 0xffffffffull, DH_E_INVALID, "dh_shard_create: too many channels");
        rc = h->sink.init(proto, (uint32_t) channels_total);
        if (rc != DH_OK) return fail(rc);
        h->sink_ready = true;
    }
    if (world > 1) {
        h->comm_in = (ncclComm_t) nccl_comm;
        // a second communicator for the gather direction: collectives of one communicator are serialised in issue
        // order, which would make scatter(k+1) wait for the kernels of step k
        ncclResult_t r = nccl().CommSplit(h->comm_in, 0, rank, &h->comm_out, nullptr);
        if (r != ncclSuccess) {
            dh::set_error("dh_shard_create: ncclCommSplit -> NCCL: %s", nccl().GetErrorString(r));
            return fail(DH_E_NCCL);
        }
    }
    if (world > 1) {
        rc = setup_ipc_scatter(h);
        if (rc != DH_OK) return fail(rc);
    }
    *out = h;
    return DH_OK;
}

size_t dh_shard_pitch(const dh_shard* h) { return h ? h->pitch : 0; }

uint32_t dh_shard_local_channels(const dh_shard* h) { return h ? h->n_local : 0; }

dh_pipe* dh_shard_pipe(dh_shard* h) { return h ? h->pipe : nullptr; }

int dh_shard_submit_device(dh_shard* h, const void* d_in, size_t pitch, size_t n, int flags, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_shard_submit_device: handle is NULL");
    DH_REQUIRE(n > 0 && n <= h->max_chunk, DH_E_INVALID, "dh_shard_submit_device: n=%zu out of range (max_chunk=%zu)", n,
               h->max_chunk);
    DH_REQUIRE(h->submitted - h->collected < 2, DH_E_STATE,
               "dh_shard_submit_device: two steps are already in flight, collect or discard one first");
    const bool scatter = (flags & DH_SHARD_SCATTER) != 0 && h->world > 1;
    const bool is_root = h->rank == h->root;
    const bool have_input = !scatter || is_root;
    DH_REQUIRE(!have_input || (d_in != nullptr && pitch == h->pitch), DH_E_INVALID,
               "dh_shard_submit_device: the input block must use the pitch dh_shard_pitch() = %zu (got %zu)", h->pitch,
               pitch);
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "dh_shard_submit_device: cannot switch to device %d", h->device);
    const uint64_t k = h->submitted;
    const int slot = (int) (k & 1);
    static const bool tracing = getenv("DH_SHARD_TRACE") != nullptr;
    auto trace_mark = [&](cudaStream_t st) {
        if (!tracing || h->trace.size() >= 6 * 64) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        h->trace.push_back(e);
    };
    const size_t row_bytes = h->pitch * h->elem;
    const char* in_local = static_cast<const char*>(d_in);

    // ---- scatter (stream s_in, communicator comm_in) --------------------------------------------------------------
    if (have_input) {
        DH_CUDA(cudaEventRecord(h->ev_user, (cudaStream_t) stream));
        DH_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_user, 0));
    }
    trace_mark(h->s_in);   // 0: scatter phase starts
    if (scatter && h->ipc_scatter) {
        // tokens: peer -> root "my slot is free", root -> peer "your rows have landed"; the rows themselves are written
        // into the peers' slots by the root's copy engines, one stream per peer
        uint32_t* tok_ready = h->d_token;               // root: [world] received; peer: the one it sends
        uint32_t* tok_done = h->d_token + h->world;
        if (is_root) {
            DH_NCCL(nccl().GroupStart());
            for (int r = 0; r < h->world; r++)
                if (r != h->root) DH_NCCL(nccl().Recv(tok_ready + r, sizeof(uint32_t), ncclInt8, r, h->comm_in, h->s_in));
            DH_NCCL(nccl().GroupEnd());
            DH_CUDA(cudaEventRecord(h->ev_ready, h->s_in));
            trace_mark(h->s_in);   // 1: all peers ready
            for (int r = 0; r < h->world; r++) {
                if (r == h->root) continue;
                DH_CUDA(cudaStreamWaitEvent(h->s_peer[r], h->ev_ready, 0));
                DH_CUDA(cudaMemcpyAsync(h->peer_slot[slot][r], in_local + h->lo_of[r] * row_bytes,
                                        (size_t) h->n_of[r] * row_bytes, cudaMemcpyDeviceToDevice, h->s_peer[r]));
                DH_CUDA(cudaEventRecord(h->ev_peer[r], h->s_peer[r]));
                DH_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_peer[r], 0));
            }
            trace_mark(h->s_in);   // 2: copies landed
            DH_NCCL(nccl().GroupStart());
            for (int r = 0; r < h->world; r++)
                if (r != h->root) DH_NCCL(nccl().Send(tok_done + r, sizeof(uint32_t), ncclInt8, r, h->comm_in, h->s_in));
            DH_NCCL(nccl().GroupEnd());
            in_local += h->lo_of[h->root] * row_bytes;
        } else {
            if (k >= 2) DH_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_consumed[slot], 0));
            trace_mark(h->s_in);   // 1: slot free
            DH_NCCL(nccl().Send(tok_ready, sizeof(uint32_t), ncclInt8, h->root, h->comm_in, h->s_in));
            trace_mark(h->s_in);   // 2: ready token sent
            DH_NCCL(nccl().Recv(tok_done, sizeof(uint32_t), ncclInt8, h->root, h->comm_in, h->s_in));
            in_local = static_cast<const char*>(h->d_slot[slot]);
        }
    } else if (scatter) {
        if (is_root) {
            DH_NCCL(nccl().GroupStart());
            for (int r = 0; r < h->world; r++) {
                if (r == h->root) continue;
                DH_NCCL(nccl().Send(in_local + h->lo_of[r] * row_bytes, (size_t) h->n_of[r] * row_bytes, ncclInt8, r,
                                    h->comm_in, h->s_in));
            }
            DH_NCCL(nccl().GroupEnd());
            in_local += h->lo_of[h->root] * row_bytes;
        } else {
            // the slot is free once the first kernel of the step that used it two submissions ago has read it
            if (k >= 2) DH_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_consumed[slot], 0));
            DH_NCCL(nccl().Recv(h->d_slot[slot], (size_t) h->n_local * row_bytes, ncclInt8, h->root, h->comm_in, h->s_in));
            in_local = static_cast<const char*>(h->d_slot[slot]);
        }
    }
    DH_CUDA(cudaEventRecord(h->ev_scattered, h->s_in));
    while (tracing && h->trace.size() % 6 != 4 && h->trace.size() < 6 * 64) trace_mark(h->s_in);   // ..3: scatter phase done

    // ---- the pipe of this rank's channels (stream s_cmp orders the input, kernels on the pipe's own streams) -------
    // the root's own rows are already in place: its kernels only wait for the caller's stream, not for the scatter
    DH_CUDA(cudaStreamWaitEvent(h->s_cmp, scatter && is_root ? h->ev_user : h->ev_scattered, 0));
    if (k >= 2) DH_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_packed[slot], 0));   // result set k & 1 is free again
    int rc = dh_decoder_select_results(h->dec, slot);
    if (rc != DH_OK) return rc;
    rc = h->fmt == DH_FMT_S16
             ? dh_pipe_process_device_s16(h->pipe, reinterpret_cast<const int16_t*>(in_local), h->pitch, n, h->s_cmp)
             : dh_pipe_process_device(h->pipe, reinterpret_cast<const float*>(in_local), h->pitch, n, h->s_cmp);
    if (rc != DH_OK) return rc;
    rc = dh_pipe_input_event(h->pipe, h->ev_consumed[slot]);
    if (rc != DH_OK) return rc;
    if (tracing && h->trace.size() % 6 == 4) {   // 4: first kernel has read the input (recorded like ev_consumed)
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) {
            dh_pipe_input_event(h->pipe, e);
            h->trace.push_back(e);
        }
    }
    DH_CUDA(cudaEventRecord(h->ev_computed, h->s_cmp));

    // ---- pack + gather (stream s_out, communicator comm_out) -------------------------------------------------------
    dh::DecoderView view;
    rc = dh::decoder_view(h->dec, slot, &view);
    if (rc != DH_OK) return rc;
    const bool fused_gather = h->ipc_gather && h->world > 1;
    uint32_t* tok_free = h->d_token + 2 * h->world;   // gather tokens live behind the scatter tokens
    uint32_t* tok_packed = h->d_token + 3 * h->world;
    // Fused form: the root first tells every peer that wire buffer `slot` is free again (its host has read or dropped
    // the step that used it, and the stream has passed that step's gather), a peer's pack kernel then stores its block
    // straight into the root's buffer over NVLink and a second token reports it.  The peers post the receive of the
    // "free" token before they wait for their kernels, so the root's send never spins for long.
    if (fused_gather) {
        if (is_root) {
            DH_NCCL(nccl().GroupStart());
            for (int r = 0; r < h->world; r++)
                if (r != h->root) DH_NCCL(nccl().Send(tok_free + r, sizeof(uint32_t), ncclInt8, r, h->comm_out, h->s_out));
            DH_NCCL(nccl().GroupEnd());
        } else {
            DH_NCCL(nccl().Recv(tok_free, sizeof(uint32_t), ncclInt8, h->root, h->comm_out, h->s_out));
        }
    }
    DH_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_computed, 0));
    rc = dh_pipe_sync(h->pipe, h->s_out);   // asynchronous pipes: the decoder kernel runs on an internal stream
    if (rc != DH_OK) return rc;
    uint8_t* my_wire = is_root ? h->d_wire[slot] + h->region_off[h->rank]
                               : (fused_gather ? static_cast<uint8_t*>(h->root_wire[slot]) + h->region_off[h->rank] : h->d_wire[slot]);
    const unsigned blocks = (h->n_local + 3) / 4 < 148u * 8u ? (h->n_local + 3) / 4 : 148u * 8u;
    pack_results_kernel<<<blocks, 128, 0, h->s_out>>>(view.counts, view.out, view.out_cap, view.ev, view.ev_cap, my_wire,
                                                       h->n_local, h->wire.w_out, h->wire.w_ev,
                                                       h->wire.off_out(h->n_local), h->wire.off_ev(h->n_local));
    DH_CUDA(cudaGetLastError());
    h->packs++;
    DH_CUDA(cudaEventRecord(h->ev_packed[slot], h->s_out));
    if (h->world > 1) {
        if (is_root) {
            DH_NCCL(nccl().GroupStart());
            for (int r = 0; r < h->world; r++) {
                if (r == h->root) continue;
                if (fused_gather)
                    DH_NCCL(nccl().Recv(tok_packed + r, sizeof(uint32_t), ncclInt8, r, h->comm_out, h->s_out));
                else
                    DH_NCCL(nccl().Recv(h->d_wire[slot] + h->region_off[r], h->wire.bytes(h->n_of[r]), ncclInt8, r,
                                        h->comm_out, h->s_out));
            }
            DH_NCCL(nccl().GroupEnd());
        } else if (fused_gather) {
            DH_NCCL(nccl().Send(tok_packed, sizeof(uint32_t), ncclInt8, h->root, h->comm_out, h->s_out));
        } else {
            DH_NCCL(nccl().Send(my_wire, h->wire.bytes(h->n_local), ncclInt8, h->root, h->comm_out, h->s_out));
        }
    }
    DH_CUDA(cudaEventRecord(h->ev_gathered[slot], h->s_out));
    if (tracing && h->trace.size() % 6 == 5) trace_mark(h->s_out);   // 5: gathered
    h->submitted++;
    return DH_OK;
}

int dh_shard_collect_step(dh_shard* h) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_shard_collect_step: handle is NULL");
    DH_REQUIRE(h->collected < h->submitted, DH_E_STATE, "dh_shard_collect_step: nothing in flight");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    const int slot = (int) (h->collected & 1);
    DH_CUDA(cudaEventSynchronize(h->ev_gathered[slot]));
    h->collected++;
    if (h->rank != h->root) return DH_OK;
    uint32_t flags = 0;
    std::vector<dh::ResultSink::Block> blocks((size_t) h->world);
    for (int r = 0; r < h->world; r++) {
        const uint32_t n = h->n_of[r];
        const uint8_t* region = h->d_wire[slot] + h->region_off[r];
        blocks[r] = {reinterpret_cast<const uint32_t*>(region), region + h->wire.off_out(n),
                     reinterpret_cast<const dh::DecEvent*>(region + h->wire.off_ev(n)), n, (uint32_t) h->lo_of[r]};
    }
    int rc = h->sink.ingest_blocks(blocks.data(), h->world, h->wire.w_out, h->wire.w_ev, h->s_back, &flags);
    if (rc != DH_OK) return rc;
    DH_REQUIRE(flags == 0, DH_E_STATE, "dh_shard_collect_step: result slots overflowed on some rank (flags 0x%x)", flags);
    return DH_OK;
}

int dh_shard_discard_step(dh_shard* h) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_shard_discard_step: handle is NULL");
    DH_REQUIRE(h->collected < h->submitted, DH_E_STATE, "dh_shard_discard_step: nothing in flight");
    h->collected++;
    return DH_OK;
}

int dh_shard_sync(dh_shard* h, void* stream) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_shard_sync: handle is NULL");
    dh::DeviceGuard guard(h->device);
    DH_REQUIRE(guard.ok, DH_E_NODEVICE, "cannot switch to CUDA device %d", h->device);
    cudaStream_t st = (cudaStream_t) stream;
    cudaStream_t all[3] = {h->s_in, h->s_cmp, h->s_out};
    for (cudaStream_t s : all) {
        DH_CUDA(cudaEventRecord(h->ev_join, s));
        DH_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    return dh_pipe_sync(h->pipe, stream);
}

int dh_shard_output(dh_shard* h, uint64_t channel, const uint8_t** data, size_t* len) {
    DH_REQUIRE(h != nullptr && h->sink_ready, DH_E_STATE, "dh_shard_output: results live on the root rank only");
    DH_REQUIRE(channel < h->channels_total, DH_E_INVALID, "dh_shard_output: channel out of range");
    if (data) *data = reinterpret_cast<const uint8_t*>(h->sink.results[channel].bytes.data());
    if (len) *len = h->sink.results[channel].bytes.size();
    return DH_OK;
}

int dh_shard_meta(dh_shard* h, uint64_t channel, const char** text, size_t* len) {
    DH_REQUIRE(h != nullptr && h->sink_ready, DH_E_STATE, "dh_shard_meta: results live on the root rank only");
    DH_REQUIRE(channel < h->channels_total, DH_E_INVALID, "dh_shard_meta: channel out of range");
    if (text) *text = h->sink.results[channel].meta.data();
    if (len) *len = h->sink.results[channel].meta.size();
    return DH_OK;
}

int dh_shard_clear(dh_shard* h) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_shard_clear: handle is NULL");
    if (h->sink_ready) h->sink.clear();
    return DH_OK;
}

int dh_shard_scatter_path(const dh_shard* h) { return h ? (h->world > 1 ? (h->ipc_scatter ? 2 : 1) : 0) : 0; }

int dh_shard_gather_path(const dh_shard* h) { return h ? (h->world > 1 ? (h->ipc_gather ? 2 : 1) : 0) : 0; }

int dh_shard_stats(dh_shard* h, uint64_t* launches, uint64_t* wire_bytes_per_step, uint64_t* d2h_bytes) {
    DH_REQUIRE(h != nullptr, DH_E_INVALID, "dh_shard_stats: handle is NULL");
    if (launches) *launches = dh_pipe_launch_count(h->pipe) + h->packs;
    if (wire_bytes_per_step) *wire_bytes_per_step = h->wire.bytes(h->n_local);
    if (d2h_bytes) *d2h_bytes = h->sink_ready ? h->sink.total_d2h : 0;
    return DH_OK;
}

void dh_shard_destroy(dh_shard* h) {
    if (!h) return;
    {
        dh::DeviceGuard guard(h->device);
        cudaDeviceSynchronize();
        for (size_t k = 0; k + 6 <= h->trace.size(); k += 6) {
            float t[6];
            for (int j = 0; j < 6; j++) cudaEventElapsedTime(&t[j], h->trace[0], h->trace[k + j]);
            fprintf(stderr, "[shard trace] rank %d step %2zu: scatter %8.3f | %8.3f | %8.3f | %8.3f  input read %8.3f  gathered %8.3f ms\n",
                    h->rank, k / 6, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (cudaEvent_t e : h->trace) cudaEventDestroy(e);
        h->trace.clear();
        if (h->comm_out && nccl().ok) nccl().CommDestroy(h->comm_out);
    }
    dh_pipe_destroy(h->pipe);
    {
        dh::DeviceGuard guard(h->device);
        for (int i = 0; i < 2; i++)
            for (void* p : h->peer_slot[i])
                if (p) cudaIpcCloseMemHandle(p);
        for (int i = 0; i < 2; i++)
            if (h->root_wire[i]) cudaIpcCloseMemHandle(h->root_wire[i]);
        for (cudaStream_t s : h->s_peer)
            if (s) cudaStreamDestroy(s);
        for (cudaEvent_t e : h->ev_peer)
            if (e) cudaEventDestroy(e);
        if (h->ev_ready) cudaEventDestroy(h->ev_ready);
        cudaFree(h->d_token);
        for (int i = 0; i < 2; i++) {
            cudaFree(h->d_slot[i]);
            cudaFree(h->d_wire[i]);
            if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
            if (h->ev_packed[i]) cudaEventDestroy(h->ev_packed[i]);
            if (h->ev_gathered[i]) cudaEventDestroy(h->ev_gathered[i]);
        }
        cudaEvent_t evs[4] = {h->ev_user, h->ev_scattered, h->ev_computed, h->ev_join};
        for (cudaEvent_t e : evs)
            if (e) cudaEventDestroy(e);
        cudaStream_t ss[4] = {h->s_in, h->s_cmp, h->s_out, h->s_back};
        for (cudaStream_t s : ss)
            if (s) cudaStreamDestroy(s);
        if (h->sink_ready) h->sink.release();
    }
    delete h;
}

}  // extern "C"
