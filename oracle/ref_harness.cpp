// ref_harness.cpp — drives the UNMODIFIED reference modules (compiled from /root/reference by
// oracle/Makefile) behind oracle_api.h.  TEST INFRASTRUCTURE ONLY; never part of the product path.
//
// What is neutralised here, and how:
//   * uninitialised heap on the hot path (src/rrc_filter/rrc_filter.cpp:9 delay line,
//     include/gfsk_demodulator.hpp:27 volume_rb, src/dmr_decoder/embedded.cpp:13): every reference TU is
//     compiled with `-include zero_heap.hpp` (malloc -> calloc) and the module objects themselves are
//     placement-constructed into zeroed storage; global operator new below is zero-filling as well.
//   * chunking: modules are always drained with `while (canProcess()) process();` (src/lib/cli.cpp:29-33).
#include <algorithm>
#include "oracle_api.h"

#include "rrc_filter.hpp"
#include "gfsk_demodulator.hpp"
#include "fsk_demodulator.hpp"
#include "digitalvoice_filter.hpp"
#include "dmr_decoder.hpp"
#include "ysf_decoder.hpp"
#include "pocsag_decoder.hpp"
#include "nxdn_decoder.hpp"
#include "dstar_decoder.hpp"
#include "meta.hpp"
#include "nxdn_decoder/trellis.hpp"
#include "nxdn_decoder/sacch.hpp"
#include "nxdn_decoder/facch1.hpp"
#include "dstar_decoder/header.hpp"

extern "C" {
#include "hamming_distance.h"
#include "hamming_7_4.h"
#include "hamming_13_9.h"
#include "hamming_15_11.h"
#include "hamming_16_11.h"
#include "quadratic_residue.h"
#include "golay_20_8.h"
#include "golay_24_12.h"
#include "bch_31_21.h"
#include "bptc_196_96.h"
#include "trellis.h"
#include "crc16.h"
#include "whitening.h"
}

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

// zero-filling global allocation for everything the reference `new`s inside this shared object
void* operator new(size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, size_t) noexcept { std::free(p); }
void operator delete[](void* p, size_t) noexcept { std::free(p); }

namespace {

    // Reader over a caller-owned array; only the first `fed` items are visible.
    template <typename T>
    class SpanReader: public Csdr::Reader<T> {
        public:
            SpanReader(const T* data): data(const_cast<T*>(data)) {}
            size_t available() override { return fed - pos; }
            T* getReadPointer() override { return data + pos; }
            void advance(size_t n) override { pos += n; }
            void feed(size_t upTo) { fed = upTo; }
            void rebase(const T* d) { data = const_cast<T*>(d); }
            size_t consumed() const { return pos; }
        private:
            T* data;
            size_t fed = 0;
            size_t pos = 0;
    };

    // Writer into a growable vector that always keeps `headroom` writable items, so that modules which
    // never check writeable() (src/dmr_decoder/dmr_phase.cpp:213-226, src/pocsag_decoder/message.cpp:22-23)
    // cannot overrun.
    template <typename T>
    class VectorWriter: public Csdr::Writer<T> {
        public:
            explicit VectorWriter(size_t headroom = 4096): headroom(headroom), buf(headroom) {}
            size_t writeable() override { return buf.size() - len; }
            T* getWritePointer() override { return buf.data() + len; }
            void advance(size_t n) override {
                len += n;
                if (buf.size() - len < headroom) buf.resize(2 * buf.size() + headroom);
            }
            const T* data() const { return buf.data(); }
            size_t size() const { return len; }
        private:
            size_t headroom;
            std::vector<T> buf;
            size_t len = 0;
    };

    class CaptureMetaWriter: public Digiham::MetaWriter {
        public:
            explicit CaptureMetaWriter(std::string* sink): Digiham::MetaWriter(), sink(sink) {}
            void sendMetaData(std::map<std::string, std::string> metadata) override {
                sink->append(serializer->serializeMetaData(metadata));
            }
        private:
            std::string* sink;
    };

    // module objects live in zeroed storage (members without initialisers read as zero)
    template <typename M, typename... Args>
    M* makeZeroed(Args... args) {
        void* mem = std::calloc(1, sizeof(M));
        return new (mem) M(args...);
    }
    template <typename M>
    void destroyZeroed(M* m) {
        m->~M();
        std::free(m);
    }

    template <typename T, typename U>
    void drain(Csdr::Module<T, U>* m) {
        while (m->canProcess()) m->process();
    }

    size_t nextFeed(size_t fed, size_t n, size_t chunk) {
        if (chunk == 0) return n;
        return fed + chunk > n ? n : fed + chunk;
    }

    Digiham::Decoder* makeDecoder(int proto, int slot_filter) {
        switch (proto) {
            case ORC_PROTO_DMR: {
                auto d = makeZeroed<Digiham::Dmr::Decoder>();
                d->setSlotFilter((unsigned char) slot_filter);
                return d;
            }
            case ORC_PROTO_YSF:
                return makeZeroed<Digiham::Ysf::Decoder>();
            case ORC_PROTO_POCSAG:
                return makeZeroed<Digiham::Pocsag::Decoder>();
            case ORC_PROTO_NXDN:
                return makeZeroed<Digiham::Nxdn::Decoder>();
            case ORC_PROTO_DSTAR:
                return makeZeroed<Digiham::DStar::Decoder>();
        }
        return nullptr;
    }

    void copyOut(const std::string& meta, char* dst, size_t cap, size_t* len) {
        size_t n = meta.size() < cap ? meta.size() : cap;
        if (dst && n) std::memcpy(dst, meta.data(), n);
        if (len) *len = n;
    }

}

extern "C" {

const char* orc_kind(void) { return "reference"; }

size_t orc_rrc(int narrow, const float* in, size_t n, size_t chunk, float* out) {
    Csdr::Module<float, float>* m = narrow
        ? (Csdr::Module<float, float>*) makeZeroed<Digiham::RrcFilter::NarrowRrcFilter>()
        : (Csdr::Module<float, float>*) makeZeroed<Digiham::RrcFilter::WideRrcFilter>();
    SpanReader<float> r(in);
    VectorWriter<float> w(n + 16);
    m->setReader(&r);
    m->setWriter(&w);
    for (size_t fed = 0; fed < n;) {
        fed = nextFeed(fed, n, chunk);
        r.feed(fed);
        drain(m);
    }
    std::memcpy(out, w.data(), w.size() * sizeof(float));
    size_t produced = w.size();
    if (narrow) destroyZeroed((Digiham::RrcFilter::NarrowRrcFilter*) m);
    else destroyZeroed((Digiham::RrcFilter::WideRrcFilter*) m);
    return produced;
}

size_t orc_demod(int four_level, unsigned sps, int invert, const float* in, size_t n, size_t chunk,
                 uint8_t* out, size_t out_cap) {
    Csdr::Module<float, unsigned char>* m = four_level
        ? (Csdr::Module<float, unsigned char>*) makeZeroed<Digiham::Fsk::GfskDemodulator>(sps)
        : (Csdr::Module<float, unsigned char>*) makeZeroed<Digiham::Fsk::FskDemodulator>(sps, invert != 0);
    SpanReader<float> r(in);
    VectorWriter<unsigned char> w;
    m->setReader(&r);
    m->setWriter(&w);
    for (size_t fed = 0; fed < n;) {
        fed = nextFeed(fed, n, chunk);
        r.feed(fed);
        drain(m);
    }
    size_t produced = w.size();
    std::memcpy(out, w.data(), produced < out_cap ? produced : out_cap);
    if (four_level) destroyZeroed((Digiham::Fsk::GfskDemodulator*) m);
    else destroyZeroed((Digiham::Fsk::FskDemodulator*) m);
    return produced;
}

size_t orc_decode(int proto, const uint8_t* sym, size_t n, size_t chunk, int slot_filter,
                  uint8_t* out, size_t out_cap, char* meta, size_t meta_cap, size_t* meta_len) {
    std::string metaText;
    Digiham::Decoder* d = makeDecoder(proto, slot_filter);
    d->setMetaWriter(new CaptureMetaWriter(&metaText));
    SpanReader<unsigned char> r(sym);
    VectorWriter<unsigned char> w;
    d->setReader(&r);
    d->setWriter(&w);
    for (size_t fed = 0; fed < n;) {
        fed = nextFeed(fed, n, chunk);
        r.feed(fed);
        drain(d);
    }
    size_t produced = w.size();
    if (out) std::memcpy(out, w.data(), produced < out_cap ? produced : out_cap);
    copyOut(metaText, meta, meta_cap, meta_len);
    d->~Decoder();
    std::free(d);
    return produced;
}

size_t orc_pipe(int proto, const float* in, size_t n, size_t chunk, int slot_filter,
                uint8_t* sym_out, size_t sym_cap, size_t* n_sym,
                uint8_t* out, size_t out_cap, char* meta, size_t meta_cap, size_t* meta_len) {
    std::string metaText;
    const bool pocsag = proto == ORC_PROTO_POCSAG;
    const bool dstar = proto == ORC_PROTO_DSTAR;
    const bool nxdn = proto == ORC_PROTO_NXDN;

    // stage buffers: each stage writes into a vector the next stage reads from
    SpanReader<float> rIn(in);
    VectorWriter<float> wFilt(n + 16);
    SpanReader<float> rFilt(wFilt.data());
    VectorWriter<unsigned char> wSym(n / 8 + 4096);
    SpanReader<unsigned char> rSym(wSym.data());
    VectorWriter<unsigned char> wOut;

    Csdr::Module<float, float>* rrc = nullptr;
    Csdr::Module<float, unsigned char>* demod;
    if (pocsag) {
        // examples/pocsag-decoder.sh:19-21: fsk_demodulator -i -s 40 | pocsag_decoder
        demod = makeZeroed<Digiham::Fsk::FskDemodulator>(40u, true);
        demod->setReader(&rIn);
    } else if (dstar) {
        // examples/dstar-decoder.sh:19-21: fsk_demodulator -s 10 | dstar_decoder
        demod = makeZeroed<Digiham::Fsk::FskDemodulator>(10u, false);
        demod->setReader(&rIn);
    } else if (nxdn) {
        // examples/nxdn48-decoder.sh:19-23: rrc_filter -n | gfsk_demodulator -s 20 | nxdn_decoder
        rrc = makeZeroed<Digiham::RrcFilter::NarrowRrcFilter>();
        rrc->setReader(&rIn);
        rrc->setWriter(&wFilt);
        demod = makeZeroed<Digiham::Fsk::GfskDemodulator>(20u);
        demod->setReader(&rFilt);
    } else {
        // examples/dmr-decoder.sh:19-23, examples/ysf-decoder.sh:19-23: rrc_filter | gfsk_demodulator
        rrc = makeZeroed<Digiham::RrcFilter::WideRrcFilter>();
        rrc->setReader(&rIn);
        rrc->setWriter(&wFilt);
        demod = makeZeroed<Digiham::Fsk::GfskDemodulator>(10u);
        demod->setReader(&rFilt);
    }
    demod->setWriter(&wSym);
    Digiham::Decoder* dec = makeDecoder(proto, slot_filter);
    dec->setMetaWriter(new CaptureMetaWriter(&metaText));
    dec->setReader(&rSym);
    dec->setWriter(&wOut);

    for (size_t fed = 0; fed < n;) {
        fed = nextFeed(fed, n, chunk);
        rIn.feed(fed);
        if (rrc) {
            drain(rrc);
            rFilt.rebase(wFilt.data());
            rFilt.feed(wFilt.size());
        }
        drain(demod);
        rSym.rebase(wSym.data());
        rSym.feed(wSym.size());
        drain(dec);
    }

    if (n_sym) *n_sym = wSym.size();
    if (sym_out) std::memcpy(sym_out, wSym.data(), wSym.size() < sym_cap ? wSym.size() : sym_cap);
    size_t produced = wOut.size();
    if (out) std::memcpy(out, wOut.data(), produced < out_cap ? produced : out_cap);
    copyOut(metaText, meta, meta_cap, meta_len);

    dec->~Decoder();
    std::free(dec);
    if (pocsag || dstar) destroyZeroed((Digiham::Fsk::FskDemodulator*) demod);
    else {
        destroyZeroed((Digiham::Fsk::GfskDemodulator*) demod);
        if (nxdn) destroyZeroed((Digiham::RrcFilter::NarrowRrcFilter*) rrc);
        else destroyZeroed((Digiham::RrcFilter::WideRrcFilter*) rrc);
    }
    return produced;
}

size_t orc_pipe_batch(int proto, const float* in, size_t nch, size_t n, size_t chunk, int slot_filter,
                      int nthreads,
                      uint8_t* sym_out, size_t sym_cap, size_t* n_sym,
                      uint8_t* out, size_t out_cap, size_t* out_len,
                      char* meta, size_t meta_cap, size_t* meta_len) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            size_t c = next.fetch_add(1);
            if (c >= nch) return;
            size_t ns = 0, ml = 0;
            size_t ol = orc_pipe(proto, in + c * n, n, chunk, slot_filter,
                                 sym_out ? sym_out + c * sym_cap : nullptr, sym_cap, &ns,
                                 out ? out + c * out_cap : nullptr, out_cap,
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

size_t orc_dvf(const int16_t* in, size_t n, size_t chunk, int16_t* out) {
    auto m = makeZeroed<Digiham::DigitalVoice::DigitalVoiceFilter>();
    Csdr::Module<short, short>* mod = m;
    SpanReader<short> r(in);
    VectorWriter<short> w(n + 16);
    mod->setReader(&r);
    mod->setWriter(&w);
    for (size_t fed = 0; fed < n;) {
        fed = nextFeed(fed, n, chunk);
        r.feed(fed);
        drain(mod);
    }
    std::memcpy(out, w.data(), w.size() * sizeof(int16_t));
    size_t produced = w.size();
    destroyZeroed(m);
    return produced;
}

int orc_fec(int code, uint32_t* word) {
    switch (code) {
        case ORC_FEC_HAMMING_7_4: { uint8_t v = (uint8_t) *word; bool ok = hamming_7_4(&v); *word = v; return ok; }
        case ORC_FEC_HAMMING_13_9: { uint16_t v = (uint16_t) *word; bool ok = hamming_13_9(&v); *word = v; return ok; }
        case ORC_FEC_HAMMING_15_11: { uint16_t v = (uint16_t) *word; bool ok = hamming_15_11(&v); *word = v; return ok; }
        case ORC_FEC_HAMMING_16_11: { uint16_t v = (uint16_t) *word; bool ok = hamming_16_11(&v); *word = v; return ok; }
        case ORC_FEC_QR_16_7: { uint16_t v = (uint16_t) *word; bool ok = quadratic_residue(&v); *word = v; return ok; }
        case ORC_FEC_GOLAY_20_8: return golay_20_8(word);
        case ORC_FEC_GOLAY_24_12: return golay_24_12(word);
        case ORC_FEC_BCH_31_21: return bch_31_21(word);
    }
    return -1;
}

uint32_t orc_fec_syndrome(int code, uint32_t word) {
    switch (code) {
        case ORC_FEC_HAMMING_7_4: { uint8_t v = (uint8_t) word; return hamming_7_4_parity(&v); }
        case ORC_FEC_HAMMING_13_9: { uint16_t v = (uint16_t) word; return hamming_13_9_parity(&v); }
        case ORC_FEC_HAMMING_15_11: { uint16_t v = (uint16_t) word; return hamming_15_11_parity(&v); }
        case ORC_FEC_HAMMING_16_11: { uint16_t v = (uint16_t) word; return hamming_16_11_parity(&v); }
        case ORC_FEC_QR_16_7: { uint16_t v = (uint16_t) word; return quadratic_residue_parity(&v); }
        case ORC_FEC_GOLAY_20_8: return golay_20_8_parity(&word);
        case ORC_FEC_GOLAY_24_12: return golay_24_12_parity(&word);
        case ORC_FEC_BCH_31_21: return bch_31_21_parity(&word);
    }
    return 0xFFFFFFFFu;
}

int orc_bptc_196_96(const uint8_t in[25], uint8_t out[12]) {
    uint8_t tmp[25];
    std::memcpy(tmp, in, 25);
    return bptc_196_96(tmp, out);
}

unsigned orc_trellis(const uint8_t* in, unsigned steps, uint8_t* out) {
    return decode_trellis(const_cast<uint8_t*>(in), (uint8_t) steps, out);
}

uint16_t orc_crc16(const uint8_t* data, int count) {
    return crc16_checksum(const_cast<uint8_t*>(data), count);
}

void orc_whitening(const uint8_t* in, uint8_t* out, unsigned nbits) {
    decode_whitening(const_cast<uint8_t*>(in), out, (uint8_t) nbits);
}

unsigned orc_hamming_distance(const uint8_t* a, const uint8_t* b, size_t n) {
    return hamming_distance(const_cast<uint8_t*>(a), const_cast<uint8_t*>(b), n);
}

unsigned orc_nxdn_trellis(const uint8_t* in, unsigned len, uint8_t* out) {
    Digiham::Nxdn::Trellis t;
    return t.decode(const_cast<uint8_t*>(in), out, len);
}

int orc_nxdn_sacch(const uint8_t in[30], uint8_t out[5]) {
    Digiham::Nxdn::Sacch* s = Digiham::Nxdn::Sacch::parse(const_cast<uint8_t*>(in));
    if (s == nullptr) return 0;
    std::memcpy(out, s->getSuperframeData() - 1, 5);
    delete s;
    return 1;
}

int orc_nxdn_facch1(const uint8_t in[72]) {
    Digiham::Nxdn::Facch1* f = Digiham::Nxdn::Facch1::parse(const_cast<uint8_t*>(in));
    if (f == nullptr) return -1;
    int t = f->getMessageType();
    delete f;
    return t;
}

int orc_dstar_header(const uint8_t in[660], char* text, size_t cap) {
    Digiham::DStar::Header* h = Digiham::DStar::Header::parseFromHeader(const_cast<uint8_t*>(in));
    if (h == nullptr) return -1;
    std::string s = h->toString();
    if (text && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(text, s.data(), n);
        text[n] = 0;
    }
    int r = h->isData() ? 1 : 0;
    delete h;
    return r;
}

}

#include "batch.inc"
