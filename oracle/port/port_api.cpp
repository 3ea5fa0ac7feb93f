// port_api.cpp — oracle_api.h on top of the CPU restatement.  TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
// `chunk` is accepted and ignored: the reference's results do not depend on how the stream is cut as long as the
// modules are drained after every chunk (asserted on the compiled reference in tests/test_oracle_cpu.py), and the
// restatement always works on the whole stream.
#include <algorithm>
#include "../oracle_api.h"
#include "port.hpp"

#include <atomic>
#include <cstring>
#include <thread>

namespace {

size_t give(const std::vector<uint8_t>& v, uint8_t* dst, size_t cap) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() < cap ? v.size() : cap);
    return v.size();
}
void giveText(const std::string& s, char* dst, size_t cap, size_t* len) {
    const size_t n = s.size() < cap ? s.size() : cap;
    if (dst && n) std::memcpy(dst, s.data(), n);
    if (len) *len = n;
}
void decode(int proto, const uint8_t* sym, size_t n, int slotFilter, port::Decoded& d) {
    switch (proto) {
        case ORC_PROTO_DMR: port::decode_dmr(sym, n, slotFilter, d); break;
        case ORC_PROTO_YSF: port::decode_ysf(sym, n, d); break;
        case ORC_PROTO_POCSAG: port::decode_pocsag(sym, n, d); break;
        case ORC_PROTO_NXDN: port::decode_nxdn(sym, n, d); break;
        case ORC_PROTO_DSTAR: port::decode_dstar(sym, n, d); break;
    }
}

}  // namespace

extern "C" {

const char* orc_kind(void) { return "port"; }

size_t orc_rrc(int narrow, const float* in, size_t n, size_t, float* out) {
    port::Rrc f(narrow != 0);
    for (size_t i = 0; i < n; i++) out[i] = f.step(in[i]);
    return n;
}

size_t orc_demod(int four_level, unsigned sps, int invert, const float* in, size_t n, size_t, uint8_t* out,
                 size_t out_cap) {
    port::Demod d(sps, four_level != 0, invert != 0);
    std::vector<uint8_t> sym;
    d.run(in, n, sym);
    return give(sym, out, out_cap);
}

size_t orc_decode(int proto, const uint8_t* sym, size_t n, size_t, int slot_filter, uint8_t* out, size_t out_cap,
                  char* meta, size_t meta_cap, size_t* meta_len) {
    port::Decoded d;
    decode(proto, sym, n, slot_filter, d);
    giveText(d.meta, meta, meta_cap, meta_len);
    return give(d.bytes, out, out_cap);
}

size_t orc_pipe(int proto, const float* in, size_t n, size_t, int slot_filter, uint8_t* sym_out, size_t sym_cap,
                size_t* n_sym, uint8_t* out, size_t out_cap, char* meta, size_t meta_cap, size_t* meta_len) {
    std::vector<uint8_t> sym;
    if (proto == ORC_PROTO_POCSAG) {
        port::Demod d(40, false, true);
        d.run(in, n, sym);
    } else if (proto == ORC_PROTO_DSTAR) {
        port::Demod d(10, false, false);
        d.run(in, n, sym);
    } else {
        const bool nxdn = proto == ORC_PROTO_NXDN;
        port::Rrc f(nxdn);
        std::vector<float> filtered(n);
        for (size_t i = 0; i < n; i++) filtered[i] = f.step(in[i]);
        port::Demod d(nxdn ? 20 : 10, true, false);
        d.run(filtered.data(), n, sym);
    }
    if (n_sym) *n_sym = sym.size();
    give(sym, sym_out, sym_cap);
    port::Decoded d;
    decode(proto, sym.data(), sym.size(), slot_filter, d);
    giveText(d.meta, meta, meta_cap, meta_len);
    return give(d.bytes, out, out_cap);
}

size_t orc_pipe_batch(int proto, const float* in, size_t nch, size_t n, size_t chunk, int slot_filter, int nthreads,
                      uint8_t* sym_out, size_t sym_cap, size_t* n_sym, uint8_t* out, size_t out_cap, size_t* out_len,
                      char* meta, size_t meta_cap, size_t* meta_len) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nch) return;
            size_t ns = 0, ml = 0;
            const size_t ol = orc_pipe(proto, in + c * n, n, chunk, slot_filter, sym_out ? sym_out + c * sym_cap : nullptr,
                                       sym_cap, &ns, out ? out + c * out_cap : nullptr, out_cap,
                                       meta ? meta + c * meta_cap : nullptr, meta_cap, &ml);
            if (n_sym) n_sym[c] = ns;
            if (out_len) out_len[c] = ol;
            if (meta_len) meta_len[c] = ml;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    return nch * n;
}

size_t orc_dvf(const int16_t* in, size_t n, size_t, int16_t* out) {
    port::Dvf f;
    for (size_t i = 0; i < n; i++) out[i] = f.step(in[i]);
    return n;
}

int orc_fec(int code, uint32_t* word) {
    if (code < 0 || code >= port::kNumCodes) return -1;
    return port::correct(code, *word) ? 1 : 0;
}

uint32_t orc_fec_syndrome(int code, uint32_t word) {
    if (code < 0 || code >= port::kNumCodes) return 0xFFFFFFFFu;
    return port::syndrome(code, word);
}

int orc_bptc_196_96(const uint8_t in[25], uint8_t out[12]) { return port::bptc_196_96(in, out) ? 1 : 0; }
unsigned orc_trellis(const uint8_t* in, unsigned steps, uint8_t* out) { return port::viterbi(in, steps, out); }
uint16_t orc_crc16(const uint8_t* data, int count) { return port::crc16(data, count); }
void orc_whitening(const uint8_t* in, uint8_t* out, unsigned nbits) { port::dewhiten(in, out, nbits); }
unsigned orc_hamming_distance(const uint8_t* a, const uint8_t* b, size_t n) { return port::hamming_distance(a, b, n); }
unsigned orc_nxdn_trellis(const uint8_t* in, unsigned len, uint8_t* out) { return port::nxdn_viterbi(in, len, out); }
int orc_nxdn_sacch(const uint8_t in[30], uint8_t out[5]) { return port::nxdn_sacch_probe(in, out); }
int orc_nxdn_facch1(const uint8_t in[72]) { return port::nxdn_facch1_probe(in); }
int orc_dstar_header(const uint8_t in[660], char* text, size_t cap) { return port::dstar_header_probe(in, text, cap); }

}

#include "../batch.inc"
