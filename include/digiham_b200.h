/*
 * digiham_b200.h — the C ABI of libdigiham_b200.so: the drop-in boundary of the B200 hot path.
 *
 * Every object below is a *bank*: N independent channels of one reference module, living on one GPU.
 * Channel c of a bank behaves bit-for-bit like one instance of the reference module it replaces, fed
 * with row c of the sample block.  All state (FIR history, timing-recovery rings, decoder phase) is
 * carried inside the bank between calls, so results do not depend on how a stream is cut into calls.
 *
 * Conventions
 *   - plain C, opaque handles, no C++/torch types in any signature;
 *   - d_* pointers are DEVICE pointers on the bank's GPU, h_* pointers are HOST pointers;
 *   - sample/symbol blocks are channel-major: element (c, t) lives at base[c * pitch + t], pitch in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); calls are
 *     asynchronous on that stream unless documented otherwise;
 *   - return value: DH_OK (0), a negative DH_E_* code, or a positive cudaError_t value;
 *     dh_last_error() returns the message of the last failure on the calling thread;
 *   - there is NO CPU fallback: without a CUDA device every create call fails.
 *
 * Reference interfaces replaced (paths relative to the reference tree, jketterl/digiham @ 410853c):
 *   dh_rrc_*      Digiham::RrcFilter::{RrcFilter,WideRrcFilter,NarrowRrcFilter}   include/rrc_filter.hpp:10-31
 *                 RrcFilter::process(float*, float*, size_t)                      src/rrc_filter/rrc_filter.cpp:16-34
 *   dh_demod_*    Digiham::Fsk::GfskDemodulator(sps)                               include/gfsk_demodulator.hpp:12-33
 *                 Digiham::Fsk::FskDemodulator(sps, invert)                        include/fsk_demodulator.hpp:12-33
 *                 canProcess()/process()              src/gfsk_demodulator/gfsk_demodulator.cpp:18-122, src/fsk_demodulator/fsk_demodulator.cpp:19-112
 *   dh_dvf_*      Digiham::DigitalVoice::DigitalVoiceFilter::process              include/digitalvoice_filter.hpp:12-19,
 *                                                                 src/digitalvoice_filter/digitalvoice_filter.cpp:6-45
 *   dh_decoder_*  Digiham::Decoder::{canProcess,process,setMetaWriter}             include/decoder.hpp:17-30, src/lib/decoder.cpp:21-47
 *                 Digiham::Dmr::Decoder::{Decoder,setSlotFilter}                   include/dmr_decoder.hpp:9-17, src/dmr_decoder/dmr_phase.cpp:35-345
 */
#ifndef DIGIHAM_B200_H
#define DIGIHAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DH_API __attribute__((visibility("default")))
#else
#define DH_API
#endif

#define DH_OK 0
#define DH_E_INVALID (-1)     /* bad argument (null handle, misaligned pointer, pitch too small ...) */
#define DH_E_NOMEM (-2)       /* host allocation failed */
#define DH_E_STATE (-3)       /* call not valid in the current state of the bank */
#define DH_E_UNSUPPORTED (-4) /* configuration not supported by the CUDA path */
#define DH_E_NODEVICE (-5)    /* no usable sm_100 device */
#define DH_E_NCCL (-6)        /* an NCCL call failed (dh_shard_*) */

DH_API const char* dh_last_error(void);
/* version of this library, formatted like Digiham::version (include/version.hpp:7) */
DH_API const char* dh_version(void);
/* number of CUDA devices visible to the library; DH_E_NODEVICE if there is none */
DH_API int dh_device_count(int* count);
/* Page-locked host blocks for the streaming host interfaces (dh_pipe_submit_host*): cudaHostAlloc, portable across
 * devices.  write_combined != 0 requests write-combined memory: faster to upload on some hosts, very slow to READ
 * from the CPU — only for blocks the host fills sequentially and never reads back. */
DH_API int dh_host_alloc(void** ptr, size_t bytes, int write_combined);
DH_API int dh_host_free(void* ptr);
/* frees the device staging the *_process_host variants cached for a bank handle (call before destroying it) */
DH_API void dh_host_scratch_release(const void* handle);

/* ------------------------------------------------------------------------------------------------------------
 * RRC FIR bank — N x Digiham::RrcFilter::RrcFilter (include/rrc_filter.hpp:10-31).
 *
 * out[c][t] = (float)((double) sum / gain), sum = ((0.0f + c0*x[t-nZeros]) + c1*x[t-nZeros+1]) + ... in that
 * order, each product and each partial sum rounded to float32, no FMA (src/rrc_filter/rrc_filter.cpp:22-34).
 * Samples before the first one ever processed read as 0 (the reference leaves them uninitialised,
 * src/rrc_filter/rrc_filter.cpp:9).
 */
typedef struct dh_rrc dh_rrc;

#define DH_RRC_WIDE 0   /* WideRrcFilter:   81 taps, gain 8.337797030  (src/rrc_filter/rrc_filter.cpp:89-115) */
#define DH_RRC_NARROW 1 /* NarrowRrcFilter: 161 taps, gain 16.67711971 (src/rrc_filter/rrc_filter.cpp:39-84)  */

DH_API int dh_rrc_create(dh_rrc** out, int device, uint32_t channels, int kind);
/* RrcFilter(nZeros, gain, coeffs) (include/rrc_filter.hpp:12): h_coeffs holds nZeros+1 taps; nZeros must be a
 * multiple of 4 and <= 1024. */
DH_API int dh_rrc_create_custom(dh_rrc** out, int device, uint32_t channels, uint32_t n_zeros, double gain,
                         const float* h_coeffs);
/* Filters n samples of every channel.  d_in/d_out must be 16-byte aligned, both pitches multiples of 4 and
 * >= n rounded up to 4 (up to 3 elements of padding after sample n-1 may be read, none is written).
 * In-place operation (d_out == d_in) is not supported. */
DH_API int dh_rrc_process(dh_rrc* h, const float* d_in, size_t in_pitch, float* d_out, size_t out_pitch, size_t n,
                   void* stream);
/* back to power-on state (zero history) */
/* The same filter fed with int16 samples: fuses `csdr convert -i s16 -o float` — the step in front of rrc_filter in
 * the reference's pipes (examples/dmr-decoder.sh:13-15), out = (float) in / SHRT_MAX — into the tile staging of the
 * FIR kernel, so the samples cross PCIe / NVLink / HBM as 2 bytes.  d_in rows are 16-byte aligned, in_pitch (in
 * int16 elements) is a multiple of 8 and >= n rounded up to 8; needs nZeros % 8 == 0 (both built-in filters).
 * float32 and int16 calls may be mixed on one bank (the carried history is float32). */
DH_API int dh_rrc_process_s16(dh_rrc* h, const int16_t* d_in, size_t in_pitch, float* d_out, size_t out_pitch, size_t n,
                       void* stream);
/* number of channels of the bank */
DH_API uint32_t dh_rrc_channels(const dh_rrc* h);
/* Host-buffer variant (synchronous): h_in/h_out are [channels][pitch] in host memory; staging is internal.
 * `channels` must equal the bank's channel count. */
DH_API int dh_rrc_process_host(dh_rrc* h, uint32_t channels, const float* h_in, size_t in_pitch, float* h_out,
                               size_t out_pitch, size_t n);
/* Tile-size policy of the FIR kernel: 0 (default) = fastest stand-alone; 1 = favour the smaller register footprint,
 * which is faster when other kernels share the SMs (set by dh_pipe_set_async). */
DH_API int dh_rrc_set_tile_preference(dh_rrc* h, int prefer_small);
DH_API int dh_rrc_reset(dh_rrc* h, void* stream);
DH_API void dh_rrc_destroy(dh_rrc* h);

/* ------------------------------------------------------------------------------------------------------------
 * FSK / GFSK demodulator bank — N x Digiham::Fsk::GfskDemodulator (four_level != 0, symbols 0..3) or
 * N x Digiham::Fsk::FskDemodulator (four_level == 0, symbols 0/1, optional inversion).
 *
 * Implements the reference's variance-minimum timing recovery and 100-symbol min/max slicer exactly
 * (src/gfsk_demodulator/gfsk_demodulator.cpp:24-122): a symbol is produced whenever more than sps + 1 samples
 * are buffered, unconsumed samples are carried inside the bank.  volume_rb, which the reference leaves
 * uninitialised (include/gfsk_demodulator.hpp:27), starts as zeros.
 */
typedef struct dh_demod dh_demod;

DH_API int dh_demod_create(dh_demod** out, int device, uint32_t channels, int four_level, uint32_t sps, int invert);
/* Zero-copy input: returns the device address (row of channel 0) and pitch where a producer such as
 * dh_rrc_process should write the next block of up to max_n samples per channel.  Passing exactly this
 * pointer/pitch to dh_demod_process skips the staging copy.  The bank alternates between two sets of rows from
 * one process call to the next (so that the producer of chunk k+1 may run while chunk k is still being
 * demodulated): query the address again before every producer call. */
DH_API int dh_demod_reserve(dh_demod* h, size_t max_n, float** d_buf, size_t* pitch);
/* upper bound of the symbols one dh_demod_process call with n samples can emit per channel */
DH_API size_t dh_demod_max_symbols(const dh_demod* h, size_t n);
/* Consumes n new samples of every channel.  Symbols of channel c are written to d_sym[c * sym_pitch ...], their
 * count for this call to d_nsym[c].  sym_pitch >= dh_demod_max_symbols(h, n). */
DH_API int dh_demod_process(dh_demod* h, const float* d_in, size_t in_pitch, size_t n, uint8_t* d_sym,
                            size_t sym_pitch, uint32_t* d_nsym, void* stream);
/* Host-buffer variant (synchronous); h_nsym receives the per-channel symbol counts of this call.
 * `channels` must equal the bank's channel count (DH_E_INVALID otherwise; likewise for the other *_process_host). */
DH_API int dh_demod_process_host(dh_demod* h, uint32_t channels, const float* h_in, size_t in_pitch, size_t n,
                                 uint8_t* h_sym, size_t sym_pitch, uint32_t* h_nsym);
DH_API uint32_t dh_demod_channels(const dh_demod* h);
/* Schedule of dh_demod_process (results are identical either way; may be changed between any two calls).
 * 0: one kernel, a lane group walks the 100-symbol blocks of its channel in order (best for thousands of channels).
 * 1: three kernels - only the variance search, whose +-1 nudge is the one sequential dependency of
 * GfskDemodulator::process (src/gfsk_demodulator/gfsk_demodulator.cpp:41-80), walks the blocks in order; the
 * window sums (:28-35) and the volume ring / slicing (:88-122) run one lane per symbol / one lane group per block
 * (lower latency for small banks).  -1 (the default of new banks; environment DH_DEMOD_SPLIT overrides it):
 * chosen per call from the bank size and the number of blocks the call spans. */
DH_API int dh_demod_set_split(dh_demod* h, int mode);
/* kernels a dh_demod_process call launches: with a fixed schedule, else those of the most recent call */
DH_API int dh_demod_kernels_per_call(const dh_demod* h);
DH_API int dh_demod_reset(dh_demod* h, void* stream);
DH_API void dh_demod_destroy(dh_demod* h);

/* ------------------------------------------------------------------------------------------------------------
 * DigitalVoiceFilter bank — N x Digiham::DigitalVoice::DigitalVoiceFilter: 10th-order Butterworth band-pass on
 * 8 kHz int16 audio (src/digitalvoice_filter/digitalvoice_filter.cpp:34-45), mixed float/double recurrence and the
 * x86 (short) truncation reproduced exactly.  In-place operation (d_out == d_in) is allowed.
 */
typedef struct dh_dvf dh_dvf;

DH_API int dh_dvf_create(dh_dvf** out, int device, uint32_t channels);
DH_API int dh_dvf_process(dh_dvf* h, const int16_t* d_in, size_t in_pitch, int16_t* d_out, size_t out_pitch,
                          size_t n, void* stream);
DH_API int dh_dvf_process_host(dh_dvf* h, uint32_t channels, const int16_t* h_in, size_t in_pitch, int16_t* h_out,
                               size_t out_pitch, size_t n);
DH_API uint32_t dh_dvf_channels(const dh_dvf* h);
DH_API int dh_dvf_reset(dh_dvf* h, void* stream);
DH_API void dh_dvf_destroy(dh_dvf* h);

/* ------------------------------------------------------------------------------------------------------------
 * Protocol decoder bank — N x Digiham::Dmr::Decoder / Ysf::Decoder / Pocsag::Decoder.
 *
 * Input: one byte per demodulated symbol (0..3 for DMR/YSF, 0/1 for POCSAG), exactly what the demodulator
 * banks emit.  Output per channel, in stream order:
 *   - the decoder's byte stream (DMR: 27-byte voice frames, src/dmr_decoder/dmr_phase.cpp:207-227);
 *   - the metadata lines a FileMetaWriter with the default StringSerializer would have written
 *     (`protocol:DMR;slot:0;...\n`, src/lib/meta.cpp:8-17,42-46).
 * Results accumulate on the device; dh_decoder_collect moves them into host buffers owned by the bank.
 */
typedef struct dh_decoder dh_decoder;

#define DH_PROTO_DMR 0
#define DH_PROTO_YSF 1
#define DH_PROTO_POCSAG 2
#define DH_PROTO_NXDN 3   /* Digiham::Nxdn::Decoder, reference include/nxdn_decoder.hpp:9-14 */
#define DH_PROTO_DSTAR 4  /* Digiham::DStar::Decoder, reference include/dstar_decoder.hpp:9-13 */

DH_API int dh_decoder_create(dh_decoder** out, int device, uint32_t channels, int proto);
/* Zero-copy input: device address (row of channel 0) and pitch where a producer such as dh_demod_process should
 * write up to max_syms symbols per channel for the next dh_decoder_process call. */
DH_API int dh_decoder_reserve(dh_decoder* h, size_t max_syms, uint8_t** d_buf, size_t* pitch);
/* Dmr::Decoder::setSlotFilter (include/dmr_decoder.hpp:12): bit 0 = slot 0, bit 1 = slot 1; channel < 0 = all.
 * Takes effect at the next dh_decoder_process call. */
DH_API int dh_decoder_set_slot_filter(dh_decoder* h, int channel, uint8_t filter);
/* Opt-in decoder modes that go BEYOND the reference (every option is off by default, and off means byte-exact
 * reference behaviour).  channel < 0 = all channels; takes effect at the next dh_decoder_process call.
 *   DH_OPT_DMR_LC_FEC  the Reed-Solomon (12,9) check of full link control words the reference leaves as a TODO
 *                      (src/dmr_decoder/lc.cpp:8-11, bptc_196_96.c:44; ETSI TS 102 361-1 B.3.6) on voice LC headers
 *                      and terminators with LC: 0 = none, 1 = verify (a word that fails is ignored), 2 = verify and
 *                      correct one octet error.  DH_E_UNSUPPORTED for other protocols. */
#define DH_OPT_DMR_LC_FEC 1
DH_API int dh_decoder_set_option(dh_decoder* h, int channel, int option, int value);
/* Consumes d_nsym[c] (<= max_nsym) new symbols of every channel c (d_nsym is a DEVICE array). */
DH_API int dh_decoder_process(dh_decoder* h, const uint8_t* d_sym, size_t sym_pitch, const uint32_t* d_nsym,
                              size_t max_nsym, void* stream);
/* Host-buffer variant (synchronous): consumes h_nsym[c] symbols of every channel from host rows and collects. */
DH_API int dh_decoder_process_host(dh_decoder* h, uint32_t channels, const uint8_t* h_sym, size_t sym_pitch,
                                   const uint32_t* h_nsym);
/* Synchronises with `stream`, copies everything produced since the last collect to the host and appends it to
 * the per-channel host buffers.  Must be called at least once every 2 process calls. */
DH_API int dh_decoder_collect(dh_decoder* h, void* stream);
/* Two device result sets exist so that a streaming caller can read step k while step k+1 is decoded:
 * dh_decoder_select_results chooses the set the next process calls append to (default 0, which is also what
 * dh_decoder_collect / dh_decoder_discard act on: they use the selected set); dh_decoder_collect_results collects a
 * given set. */
DH_API int dh_decoder_select_results(dh_decoder* h, int set);
DH_API int dh_decoder_collect_results(dh_decoder* h, int set, void* stream);
/* Host views of one channel's accumulated results; valid until the next collect / clear / destroy. */
DH_API int dh_decoder_output(dh_decoder* h, uint32_t channel, const uint8_t** data, size_t* len);
DH_API int dh_decoder_meta(dh_decoder* h, uint32_t channel, const char** text, size_t* len);
/* The same metadata updates as key/value records, for callers that apply their own Digiham::Serializer
 * (include/meta.hpp:10-19): per update u16 pair count, then per pair u16 key length, key, u16 value length, value
 * (little endian, keys in std::map order). */
DH_API int dh_decoder_meta_kv(dh_decoder* h, uint32_t channel, const uint8_t** data, size_t* len);
/* The key/value records cost host time per update, so they are only kept by default for banks of <= 64 channels (the
 * facade modules use one-channel banks); this switches them on or off for any bank from the next collect on. */
DH_API int dh_decoder_set_meta_kv(dh_decoder* h, int enable);
/* totals over all channels since creation */
DH_API int dh_decoder_totals(dh_decoder* h, uint64_t* out_bytes, uint64_t* meta_bytes);
/* events replayed and bytes copied device->host by collect since creation */
DH_API int dh_decoder_stats(dh_decoder* h, uint64_t* events, uint64_t* d2h_bytes);
/* Drops the results pending ON THE DEVICE without copying them (asynchronous counter reset).  For callers that
 * consume the device buffers themselves or only measure the kernels. */
DH_API int dh_decoder_discard(dh_decoder* h, void* stream);
/* Host-only helper (no GPU involved): replays n_events 16-byte decoder event records of ONE channel, in order,
 * through a fresh metadata collector of the given protocol and returns the lines it would have written.
 * Record layout: {u8 kind, u8 slot, u8 a, u8 b, u8 data[12]}; DMR kinds: 1 slot reset, 2 set sync a (+ soft reset
 * if b), 3 soft reset, 4 link control (9 LC bytes in data), 5 talker-alias collector reset. */
DH_API int dh_meta_replay(int proto, const void* events, uint32_t n_events, char* out, size_t cap, size_t* len);
DH_API uint32_t dh_decoder_channels(const dh_decoder* h);
/* drops the accumulated host results */
DH_API int dh_decoder_clear(dh_decoder* h);
DH_API void dh_decoder_destroy(dh_decoder* h);

/* ------------------------------------------------------------------------------------------------------------
 * Whole pipe of one protocol for N channels, wired like the reference's example scripts:
 *   DH_PROTO_DMR / DH_PROTO_YSF : rrc_filter | gfsk_demodulator (sps 10) | {dmr,ysf}_decoder
 *                                 (examples/dmr-decoder.sh:19-23, examples/ysf-decoder.sh:19-23)
 *   DH_PROTO_POCSAG             : fsk_demodulator -i -s 40 | pocsag_decoder   (examples/pocsag-decoder.sh:19-21)
 * Input is what the first module of those pipes reads: 48 kHz float32 FM-discriminator audio, one row per
 * channel.  The stages hand their blocks over in device memory without copies.
 */
typedef struct dh_pipe dh_pipe;

/* max_chunk: the largest n a process call will pass */
DH_API int dh_pipe_create(dh_pipe** out, int device, uint32_t channels, int proto, size_t max_chunk);
/* n samples per channel from DEVICE memory (16-byte aligned, pitch % 4 == 0, pitch >= n rounded up to 4) */
DH_API int dh_pipe_process_device(dh_pipe* h, const float* d_in, size_t in_pitch, size_t n, void* stream);
/* int16 input (`csdr convert -i s16 -o float` fused into the RRC stage, see dh_rrc_process_s16): rows 16-byte
 * aligned, pitch % 8 == 0, pitch >= n rounded up to 8.  DH_E_UNSUPPORTED for the pipes without an RRC stage. */
DH_API int dh_pipe_process_device_s16(dh_pipe* h, const int16_t* d_in, size_t in_pitch, size_t n, void* stream);
/* Records `event` (a cudaEvent_t) at the point where the most recent dh_pipe_process_device* call has finished
 * reading its input block, so that a producer can reuse the block without waiting for the whole pipe. */
DH_API int dh_pipe_input_event(dh_pipe* h, void* event);
DH_API uint32_t dh_pipe_channels(const dh_pipe* h);
/* n samples per channel from HOST memory (pinned memory makes the copy asynchronous); the host-to-device copy
 * is part of the call.  A pitch of dh_pipe_host_pitch() allows one contiguous transfer. */
DH_API int dh_pipe_process_host(dh_pipe* h, const float* h_in, size_t in_pitch, size_t n, void* stream);
DH_API int dh_pipe_process_host_s16(dh_pipe* h, const int16_t* h_in, size_t in_pitch, size_t n, void* stream);
DH_API size_t dh_pipe_host_pitch(const dh_pipe* h);
DH_API size_t dh_pipe_host_pitch_s16(const dh_pipe* h);
/* Streaming host interface: up to two steps in flight.  dh_pipe_submit_host starts the asynchronous upload of a
 * block from PINNED host memory (which must stay untouched until the step is collected) followed by the three
 * kernels on internal streams and returns at once; dh_pipe_collect_step waits for the OLDEST step in flight, reads
 * its frames / metadata back and appends them to the per-channel host buffers.  The upload of step k+1 overlaps
 * kernels, read-back and metadata replay of step k.  Do not mix with dh_pipe_process_* on the same pipe while steps
 * are in flight.  DH_E_STATE when two steps are already in flight / nothing is in flight. */
DH_API int dh_pipe_submit_host(dh_pipe* h, const float* h_in, size_t in_pitch, size_t n);
DH_API int dh_pipe_submit_host_s16(dh_pipe* h, const int16_t* h_in, size_t in_pitch, size_t n);
DH_API int dh_pipe_collect_step(dh_pipe* h);
/* same contract as dh_decoder_collect */
DH_API int dh_pipe_collect(dh_pipe* h, void* stream);
/* the decoder bank of the pipe: use dh_decoder_output / _meta / _totals / _clear / _set_slot_filter on it */
DH_API dh_decoder* dh_pipe_decoder(dh_pipe* h);
/* the demodulator bank of the pipe (dh_demod_set_split / dh_demod_kernels_per_call on it) */
DH_API dh_demod* dh_pipe_demod(dh_pipe* h);
/* device views of the symbols the demodulator produced in the LAST process call (parity artefact) */
DH_API int dh_pipe_last_symbols(dh_pipe* h, const uint8_t** d_sym, size_t* sym_pitch, const uint32_t** d_nsym);
/* Software pipelining inside one process call: the chunk is cut into sub-chunks of `sub_chunk` samples (multiple
 * of 4) and K1 of sub-chunk c+1 runs concurrently with K2 + decoder of sub-chunk c on two internal streams; the
 * caller's stream is joined before the call returns control of it.  0 (the default) runs the three kernels back to
 * back on the caller's stream.  Results are identical either way. */
DH_API int dh_pipe_set_sub_chunk(dh_pipe* h, size_t sub_chunk);
/* Software pipelining ACROSS dh_pipe_process_device calls (off by default).  When enabled, a call only orders
 * itself after the work already enqueued on `stream` (the producer of d_in) and returns with its three kernels
 * enqueued on internal streams: the RRC kernel of call i+1 overlaps the demodulator + decoder kernels of call i
 * (the FIR is issue-bound, the other two are latency-bound walks).  Results are bit-identical.  The caller joins
 * with dh_pipe_sync (makes `stream` wait for everything enqueued so far; also required before d_in is
 * overwritten) or dh_pipe_collect (which syncs first); dh_pipe_discard drops device-side results in pipeline
 * order.  In the reference the same overlap exists between the processes of a shell pipe
 * (examples/dmr-decoder.sh:19-23). */
DH_API int dh_pipe_set_async(dh_pipe* h, int enable, void* stream);
DH_API int dh_pipe_sync(dh_pipe* h, void* stream);
DH_API int dh_pipe_discard(dh_pipe* h, void* stream);
/* Per-stage device timing: when enabled every process call records CUDA events on the caller's stream around
 * K1 (RRC), K2 (demodulator) and the decoder kernel, each on the stream the kernel runs on.
 * dh_pipe_stage_times waits for the recorded work, returns the summed milliseconds per stage since the previous
 * query and the number of (sub-)chunks they cover.  With pipelining the stages overlap, so the three sums can
 * exceed the wall time of the call. */
DH_API int dh_pipe_set_profiling(dh_pipe* h, int enable);
DH_API int dh_pipe_stage_times(dh_pipe* h, double ms[3], uint64_t* calls);
/* kernels launched by this pipe since creation */
DH_API uint64_t dh_pipe_launch_count(const dh_pipe* h);
/* synchronous host copy of one channel's symbols of the last process call (test / debug helper) */
DH_API int dh_pipe_read_symbols(dh_pipe* h, uint32_t channel, uint8_t* h_buf, size_t cap, size_t* count);
DH_API void dh_pipe_destroy(dh_pipe* h);

/* ------------------------------------------------------------------------------------------------------------
 * Per-channel state as opaque blobs (checkpoint / migration of running channels).  In the reference the state of
 * a channel is the member data of its module instances (delay line: src/rrc_filter/rrc_filter.cpp:5-10; rings and
 * offsets: include/gfsk_demodulator.hpp:20-33; phase + collectors: include/decoder.hpp:24-29); here it lives in HBM
 * (plus the host-side metadata collectors of a decoder bank).  export writes header + state into a HOST buffer of
 * at least *_state_size bytes (synchronises `stream`); import loads a blob into a bank of the SAME configuration
 * (kind, channel count, taps / sps / protocol: DH_E_INVALID otherwise), after which the bank continues the stream
 * exactly where the exporting bank stopped.  Decoded results that were not collected yet are not part of the state.
 * The dh_pipe_* forms bundle the blobs of the pipe's stages (the pipe must have no step in flight).
 */
DH_API int dh_rrc_state_size(const dh_rrc* h, size_t* bytes);
DH_API int dh_rrc_state_export(dh_rrc* h, void* h_buf, size_t cap, size_t* written, void* stream);
DH_API int dh_rrc_state_import(dh_rrc* h, const void* h_buf, size_t bytes, void* stream);
DH_API int dh_demod_state_size(const dh_demod* h, size_t* bytes);
DH_API int dh_demod_state_export(dh_demod* h, void* h_buf, size_t cap, size_t* written, void* stream);
DH_API int dh_demod_state_import(dh_demod* h, const void* h_buf, size_t bytes, void* stream);
DH_API int dh_decoder_state_size(const dh_decoder* h, size_t* bytes);
DH_API int dh_decoder_state_export(dh_decoder* h, void* h_buf, size_t cap, size_t* written, void* stream);
DH_API int dh_decoder_state_import(dh_decoder* h, const void* h_buf, size_t bytes, void* stream);
DH_API int dh_pipe_state_size(const dh_pipe* h, size_t* bytes);
DH_API int dh_pipe_state_export(dh_pipe* h, void* h_buf, size_t cap, size_t* written, void* stream);
DH_API int dh_pipe_state_import(dh_pipe* h, const void* h_buf, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * One pipe sharded over the GPUs of a node: one process (rank) per GPU, channels split into contiguous ranges
 * (rank r of R owns [r*N/R, (r+1)*N/R), the first N % R ranks one channel more).  In the reference the same
 * partitioning is "one shell pipe per channel" (examples/dmr-decoder.sh:12-23): channels never interact, so the only
 * exchange steps are moving samples in and decoded frames out.  NCCL (>= 2.18, resolved at run time from the copy
 * loaded in the process) carries both over NVLink:
 *   scatter  the root (ingest) rank holds the block of ALL channels, every peer receives its rows — written into
 *            the peers' input slots by the root's copy engines over NVLink (slots mapped through CUDA IPC, NCCL
 *            only carries 4-byte ordering tokens), or sent with ncclSend / ncclRecv when that mapping is unavailable;
 *   gather   every rank packs its decoder results of the step (frames, metadata events, counts) into one wire block —
 *            stored by the pack kernel directly into the root's memory over NVLink (CUDA IPC), or sent with
 *            ncclSend / ncclRecv — and the root exposes all channels through dh_shard_output / dh_shard_meta.
 * Steps are pipelined: scatter of step k+1, the kernels of step k and the gather of step k-1 overlap (three streams,
 * two communicators); up to two steps may be in flight.  Every call below is COLLECTIVE: all ranks make the same
 * sequence of create / submit / collect / discard calls with the same n and flags.
 */
typedef struct dh_shard dh_shard;

#define DH_FMT_F32 0 /* float32 samples (what rrc_filter / fsk_demodulator read) */
#define DH_FMT_S16 1 /* int16 samples; `csdr convert -i s16 -o float` is fused into the RRC stage */
#define DH_SHARD_SCATTER 1 /* submit flag: d_in on the root covers all channels and is scattered */

/* Bootstrap helpers for hosts without a communicator of their own (thin wrappers over ncclGetUniqueId /
 * ncclCommInitRank): rank 0 creates the 128-byte id and hands it to the other ranks by any means. */
DH_API int dh_shard_unique_id(uint8_t id[128]);
DH_API int dh_shard_comm_init(void** comm, const uint8_t id[128], int rank, int world, int device);
DH_API int dh_shard_comm_destroy(void* comm);
/* the channel range [*lo, *hi) of a rank; host-only */
DH_API int dh_shard_channel_range(uint64_t channels_total, int world, int rank, uint64_t* lo, uint64_t* hi);
/* wire block of `channels` channels for steps of up to max_chunk samples: bytes / event records per channel slot and
 * the size of the whole block; host-only */
DH_API int dh_shard_wire_layout(int proto, size_t max_chunk, uint32_t channels, uint32_t* slot_bytes,
                                uint32_t* slot_events, size_t* block_bytes);
/* nccl_comm: an ncclComm_t spanning the `world` ranks (this process = `rank`, on GPU `device`); it stays owned by the
 * caller and must outlive the shard.  root: the ingest + gathering rank.  world == 1 needs no communicator. */
DH_API int dh_shard_create(dh_shard** out, void* nccl_comm, int rank, int world, int root, int device,
                           uint64_t channels_total, int proto, size_t max_chunk, int sample_format);
/* row pitch (in samples) every input block must use */
DH_API size_t dh_shard_pitch(const dh_shard* h);
DH_API uint32_t dh_shard_local_channels(const dh_shard* h);
/* the pipe of this rank's channels (e.g. for dh_pipe_last_symbols / dh_decoder_set_slot_filter on its decoder) */
DH_API dh_pipe* dh_shard_pipe(dh_shard* h);
/* One step of n samples per channel, enqueued asynchronously.  With DH_SHARD_SCATTER the root passes the DEVICE block
 * [channels_total][pitch] (produced on `stream`) and the other ranks pass NULL; without it every rank passes the
 * block of its own channels.  The block must stay untouched until dh_shard_sync / the step is collected.
 * DH_E_STATE when two steps are already in flight. */
DH_API int dh_shard_submit_device(dh_shard* h, const void* d_in, size_t pitch, size_t n, int flags, void* stream);
/* Waits for the oldest step in flight.  On the root: reads the gathered wire blocks back and appends frames and
 * metadata lines of ALL channels to the per-channel host buffers. */
DH_API int dh_shard_collect_step(dh_shard* h);
/* Retires the oldest step in flight without reading it back and without waiting (device-side throughput runs). */
DH_API int dh_shard_discard_step(dh_shard* h);
/* makes `stream` wait for everything enqueued so far */
DH_API int dh_shard_sync(dh_shard* h, void* stream);
/* root only: accumulated results of GLOBAL channel c; valid until the next collect / clear / destroy */
DH_API int dh_shard_output(dh_shard* h, uint64_t channel, const uint8_t** data, size_t* len);
DH_API int dh_shard_meta(dh_shard* h, uint64_t channel, const char** text, size_t* len);
DH_API int dh_shard_clear(dh_shard* h);
/* how the rows travel in a scattering submit: 0 = single rank, 1 = NCCL send / recv, 2 = the root's copy engines write
 * into the peers' input slots mapped through CUDA IPC (chosen at create; DH_SHARD_NO_IPC=1 in the environment forces 1) */
DH_API int dh_shard_scatter_path(const dh_shard* h);
/* how the wire blocks travel to the root: 0 = single rank, 1 = ncclSend / ncclRecv, 2 = every peer's pack kernel stores
 * its block straight into the root's wire buffer mapped through CUDA IPC (pack + gather in one kernel, NCCL carries two
 * 4-byte tokens per peer and step); DH_SHARD_NO_IPC=1 or DH_SHARD_NO_IPC_GATHER=1 force 1 */
DH_API int dh_shard_gather_path(const dh_shard* h);
/* kernels launched by this rank (pipe + pack), bytes of its wire block per step, bytes read back by collect (root) */
DH_API int dh_shard_stats(dh_shard* h, uint64_t* launches, uint64_t* wire_bytes_per_step, uint64_t* d2h_bytes);
DH_API void dh_shard_destroy(dh_shard* h);

/* ------------------------------------------------------------------------------------------------------------
 * Test hooks: the DEVICE implementations of the FEC primitives, one word / block / trellis input per element, for
 * exhaustive comparison with the reference's C functions.  Host buffers, current device, synchronous.
 *   dh_test_fec  code 0 hamming_7_4 (src/dmr_decoder/hamming_7_4.c:40-72), 1 hamming_13_9 (hamming_13_9.c:52-84),
 *                2 hamming_15_11 (hamming_15_11.c:56-88), 3 hamming_16_11 (hamming_16_11.c:61-93),
 *                4 quadratic_residue (quadratic_residue.c:302-335), 5 golay_20_8 (golay_20_8.c:1403-1435),
 *                6 golay_24_12 (src/ysf_decoder/golay_24_12.c:2383-2415), 7 bch_31_21
 *                (src/pocsag_decoder/bch_31_21.c:521-561): words are corrected in place, ok[i] = return value.
 *   dh_test_bptc bptc_196_96 (src/dmr_decoder/bptc_196_96.c:5-59): payload [n][25] bytes -> out [n][12], ok[n].
 *   dh_test_viterbi variant 0 / 1: decode_trellis with 100 / 180 steps (src/ysf_decoder/trellis.c:32-109);
 *                variant 2 / 3: Nxdn::Trellis::decode with 36 / 96 steps (src/nxdn_decoder/trellis.cpp:29-101).
 *                dibits [n][steps], one received dibit per byte; words [n][(steps + 31) / 32] decoded bits MSB first;
 *                metric [n] the winning path metric (uint8 / uint16 arithmetic like the references).
 */
DH_API int dh_test_fec(int code, uint32_t* h_words, uint8_t* h_ok, uint32_t n);
DH_API int dh_test_bptc(const uint8_t* h_payload, uint8_t* h_out, uint8_t* h_ok, uint32_t n);
DH_API int dh_test_viterbi(int variant, const uint8_t* h_dibits, uint32_t n, uint32_t* h_words, uint32_t* h_metric);

#ifdef __cplusplus
}
#endif

#endif /* DIGIHAM_B200_H */
