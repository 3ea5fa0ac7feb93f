/*
 * oracle_api.h — C interface shared by BOTH CPU checkers of this repository.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (digiham_b200/, include/) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and there only as the checker.
 *
 * Two shared objects export exactly these symbols:
 *   oracle/_ref/libdigiham_ref.so   the UNMODIFIED reference sources from /root/reference compiled
 *                                   against oracle/csdr_shim (recipe: oracle/Makefile, driver:
 *                                   oracle/ref_harness.cpp)                    -> kind "reference"
 *   oracle/liboracle_port.so        the independent C++ restatement in oracle/port/ -> kind "port"
 *
 * Every call builds fresh module instances, i.e. one call == one channel from power-on, with all
 * state that the reference leaves uninitialised forced to zero.
 * `chunk` emulates streaming: the input is made visible to the module chunk items at a time and the
 * module is drained with `while (canProcess()) process();` after every chunk (src/lib/cli.cpp:29-33).
 * chunk == 0 means "everything at once".
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PROTO_DMR = 0, ORC_PROTO_YSF = 1, ORC_PROTO_POCSAG = 2, ORC_PROTO_NXDN = 3, ORC_PROTO_DSTAR = 4 };

enum {
    ORC_FEC_HAMMING_7_4 = 0,
    ORC_FEC_HAMMING_13_9 = 1,
    ORC_FEC_HAMMING_15_11 = 2,
    ORC_FEC_HAMMING_16_11 = 3,
    ORC_FEC_QR_16_7 = 4,
    ORC_FEC_GOLAY_20_8 = 5,
    ORC_FEC_GOLAY_24_12 = 6,
    ORC_FEC_BCH_31_21 = 7
};

/* "reference" or "port" */
const char* orc_kind(void);

/* RrcFilter (include/rrc_filter.hpp:10-31). narrow: 0 = WideRrcFilter (81 taps), 1 = NarrowRrcFilter (161). */
size_t orc_rrc(int narrow, const float* in, size_t n, size_t chunk, float* out);

/* GfskDemodulator(sps) when four_level != 0, else FskDemodulator(sps, invert).  Returns #symbols. */
size_t orc_demod(int four_level, unsigned sps, int invert, const float* in, size_t n, size_t chunk,
                 uint8_t* out, size_t out_cap);

/* Dmr:: / Ysf:: / Pocsag:: / Nxdn:: / DStar::Decoder on a symbol stream.  Returns #bytes written to out.
 * meta receives the concatenated StringSerializer lines of a MetaWriter attached before the first
 * symbol; *meta_len the byte count (truncated at meta_cap).  slot_filter only applies to DMR. */
size_t orc_decode(int proto, const uint8_t* sym, size_t n, size_t chunk, int slot_filter,
                  uint8_t* out, size_t out_cap, char* meta, size_t meta_cap, size_t* meta_len);

/* The whole pipe of one channel, wired like examples/{dmr,ysf,pocsag}-decoder.sh:
 *   DMR/YSF: WideRrcFilter -> GfskDemodulator(10) -> decoder;  POCSAG: FskDemodulator(40, true) -> decoder;
 *   NXDN: NarrowRrcFilter -> GfskDemodulator(20) -> decoder (examples/nxdn48-decoder.sh:19-23);
 *   D-Star: FskDemodulator(10) -> decoder (examples/dstar-decoder.sh:19-21).
 * sym_out (nullable) receives the demodulator output. */
size_t orc_pipe(int proto, const float* in, size_t n, size_t chunk, int slot_filter,
                uint8_t* sym_out, size_t sym_cap, size_t* n_sym,
                uint8_t* out, size_t out_cap, char* meta, size_t meta_cap, size_t* meta_len);

/* nch independent channels of n samples each (row-major [nch][n]) on nthreads host threads, one
 * channel per thread at a time.  out: [nch][out_cap], out_len[nch]; sym_out nullable [nch][sym_cap];
 * meta nullable [nch][meta_cap].  Returns the total number of input samples consumed (nch * n). */
size_t orc_pipe_batch(int proto, const float* in, size_t nch, size_t n, size_t chunk, int slot_filter,
                      int nthreads,
                      uint8_t* sym_out, size_t sym_cap, size_t* n_sym,
                      uint8_t* out, size_t out_cap, size_t* out_len,
                      char* meta, size_t meta_cap, size_t* meta_len);

/* DigitalVoiceFilter (include/digitalvoice_filter.hpp:12-19). */
size_t orc_dvf(const int16_t* in, size_t n, size_t chunk, int16_t* out);

/* Block codes: corrects *word in place, returns 1 on success and 0 when uncorrectable. */
int orc_fec(int code, uint32_t* word);
/* Syndrome ("*_parity") of the same codes. */
uint32_t orc_fec_syndrome(int code, uint32_t word);

int orc_bptc_196_96(const uint8_t in[25], uint8_t out[12]);
/* decode_trellis (src/ysf_decoder/trellis.c:32): steps dibits in, (steps+7)/8 bytes out, returns best metric. */
unsigned orc_trellis(const uint8_t* in, unsigned steps, uint8_t* out);
uint16_t orc_crc16(const uint8_t* data, int count);
void orc_whitening(const uint8_t* in, uint8_t* out, unsigned nbits);
unsigned orc_hamming_distance(const uint8_t* a, const uint8_t* b, size_t n);

/* Nxdn::Trellis::decode (src/nxdn_decoder/trellis.cpp:29-101): len input BITS (len/2 dibits packed 4 per byte,
 * MSB first), (len+15)/16 bytes out, returns the best metric. */
unsigned orc_nxdn_trellis(const uint8_t* in, unsigned len, uint8_t* out);
/* Nxdn::Sacch::parse on 30 descrambled dibits (sacch.cpp:24-43): 1 + writes the 5 decoded bytes, or 0. */
int orc_nxdn_sacch(const uint8_t in[30], uint8_t out[5]);
/* Nxdn::Facch1::parse on 72 descrambled dibits (facch1.cpp:8-26): returns the message type, or -1. */
int orc_nxdn_facch1(const uint8_t in[72]);
/* DStar::Header::parseFromHeader on 660 raw bits (header.cpp:23-48): -1 when rejected, else isData() and
 * Header::toString() in text (NUL-terminated, truncated at cap). */
int orc_dstar_header(const uint8_t in[660], char* text, size_t cap);

/* Batched forms (batch.inc): loops over the single-item functions above on nthreads host threads.
 * orc_trellis_batch: in [n][in_stride] packed dibits, out [n][out_stride]; nxdn != 0 runs orc_nxdn_trellis. */
void orc_fec_batch(int code, uint32_t* words, uint8_t* ok, size_t n, int nthreads);
void orc_bptc_batch(const uint8_t* in, uint8_t* out, uint8_t* ok, size_t n, int nthreads);
void orc_trellis_batch(int nxdn, const uint8_t* in, size_t in_stride, unsigned steps, uint8_t* out, size_t out_stride,
                       unsigned* metric, size_t n, int nthreads);

#ifdef __cplusplus
}
#endif
