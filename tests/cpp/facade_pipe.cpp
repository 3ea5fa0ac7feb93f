// facade_pipe.cpp — BASELINE config 1 ("1 ch synthetic 48 kHz 4-FSK -> rrc_filter -> gfsk_demodulator ...,
// plumbing"): drives the header-compatible facade modules of include/*.hpp exactly like the reference's CLI
// template does (reference src/lib/cli.cpp:19-37,101-106: 128-item reads into a 1024-item ring, then
// `while (module->canProcess()) module->process();`), chained through csdr ring buffers.
// The csdr headers come from oracle/csdr_shim (libcsdr is not installed in this image); test code only.
//
// usage: facade_pipe <proto: dmr|ysf|pocsag|nxdn|dstar|rrc|dvf> <in file> <out prefix>
#include <csdr/ringbuffer.hpp>

#include "rrc_filter.hpp"
#include "gfsk_demodulator.hpp"
#include "fsk_demodulator.hpp"
#include "digitalvoice_filter.hpp"
#include "dmr_decoder.hpp"
#include "ysf_decoder.hpp"
#include "pocsag_decoder.hpp"
#include "nxdn_decoder.hpp"
#include "dstar_decoder.hpp"
#include "version.hpp"

#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

template <typename T>
class FileWriter: public Csdr::Writer<T> {
    public:
        explicit FileWriter(const std::string& path): out(path, std::ios::binary), buf(4096) {}
        size_t writeable() override { return buf.size(); }
        T* getWritePointer() override { return buf.data(); }
        void advance(size_t n) override {
            out.write((const char*) buf.data(), n * sizeof(T));
            out.flush();   // the writers are leaked like in the reference CLI (src/lib/cli.cpp:26-27,35)
        }
    private:
        std::ofstream out;
        std::vector<T> buf;
};

template <typename T>
std::vector<T> readAll(const char* path) {
    std::ifstream in(path, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<T> v(raw.size() / sizeof(T));
    std::memcpy(v.data(), raw.data(), v.size() * sizeof(T));
    return v;
}

// src/lib/cli.cpp:101-106
template <typename T>
bool feed(Csdr::Ringbuffer<T>* rb, const std::vector<T>& data, size_t& pos) {
    if (pos >= data.size()) return false;
    size_t n = std::min<size_t>(128, data.size() - pos);
    n = std::min(n, rb->writeable());
    std::memcpy(rb->getWritePointer(), data.data() + pos, n * sizeof(T));
    rb->advance(n);
    pos += n;
    return true;
}

// a serializer whose output does not depend on separators inside the values: key=<len>:<value> per field, then 0x1e
class LengthPrefixedSerializer: public Digiham::Serializer {
    public:
        std::string serializeMetaData(std::map<std::string, std::string> metadata) override {
            std::string out;
            for (const auto& kv : metadata) out += kv.first + "=" + std::to_string(kv.second.size()) + ":" + kv.second;
            out += '\x1e';
            return out;
        }
};

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string proto = argv[1], prefix = argv[3];
    std::cerr << "digiham-b200 " << Digiham::version << "\n";
    try {
        if (proto == "rrc" || proto == "dvf") {
            if (proto == "rrc") {
                auto data = readAll<float>(argv[2]);
                Csdr::Ringbuffer<float> ring(1024);
                Csdr::Module<float, float>* m = new Digiham::RrcFilter::WideRrcFilter();
                m->setReader(new Csdr::RingbufferReader<float>(&ring));
                m->setWriter(new FileWriter<float>(prefix + ".out"));
                size_t pos = 0;
                while (feed(&ring, data, pos)) while (m->canProcess()) m->process();
                delete m;
            } else {
                auto data = readAll<short>(argv[2]);
                Csdr::Ringbuffer<short> ring(1024);
                Csdr::Module<short, short>* m = new Digiham::DigitalVoice::DigitalVoiceFilter();
                m->setReader(new Csdr::RingbufferReader<short>(&ring));
                m->setWriter(new FileWriter<short>(prefix + ".out"));
                size_t pos = 0;
                while (feed(&ring, data, pos)) while (m->canProcess()) m->process();
                delete m;
            }
            return 0;
        }
        auto data = readAll<float>(argv[2]);
        Csdr::Ringbuffer<float> in(1024), filt(1024);
        Csdr::Ringbuffer<unsigned char> syms(1024);
        Csdr::Module<float, float>* rrc = nullptr;
        Csdr::Module<float, unsigned char>* demod;
        Digiham::Decoder* dec;
        if (proto == "pocsag" || proto == "pocsag_custom") {
            demod = new Digiham::Fsk::FskDemodulator(40, true);
            demod->setReader(new Csdr::RingbufferReader<float>(&in));
            // reference include/pocsag_decoder.hpp:12: Decoder(Serializer*) — a caller-supplied serializer receives
            // the structured {address, message} map of every message
            if (proto == "pocsag_custom") dec = new Digiham::Pocsag::Decoder(new LengthPrefixedSerializer());
            else dec = new Digiham::Pocsag::Decoder();
        } else if (proto == "dstar") {
            // examples/dstar-decoder.sh:19-21
            demod = new Digiham::Fsk::FskDemodulator(10);
            demod->setReader(new Csdr::RingbufferReader<float>(&in));
            dec = new Digiham::DStar::Decoder();
        } else {
            const bool nxdn = proto == "nxdn";   // examples/nxdn48-decoder.sh:19-23: rrc_filter -n | gfsk -s 20
            if (nxdn) rrc = new Digiham::RrcFilter::NarrowRrcFilter();
            else rrc = new Digiham::RrcFilter::WideRrcFilter();
            rrc->setReader(new Csdr::RingbufferReader<float>(&in));
            rrc->setWriter(&filt);
            demod = new Digiham::Fsk::GfskDemodulator(nxdn ? 20 : 10);
            demod->setReader(new Csdr::RingbufferReader<float>(&filt));
            if (proto == "dmr") dec = new Digiham::Dmr::Decoder();
            else if (nxdn) dec = new Digiham::Nxdn::Decoder();
            else dec = new Digiham::Ysf::Decoder();
        }
        demod->setWriter(&syms);
        dec->setReader(new Csdr::RingbufferReader<unsigned char>(&syms));
        dec->setWriter(new FileWriter<unsigned char>(prefix + ".out"));
        dec->setMetaWriter(new Digiham::FileMetaWriter(fopen((prefix + ".meta").c_str(), "wb")));
        size_t pos = 0;
        while (feed(&in, data, pos)) {
            bool progress = true;
            while (progress) {
                progress = false;
                if (rrc) while (rrc->canProcess()) { rrc->process(); progress = true; }
                while (demod->canProcess()) { demod->process(); progress = true; }
                while (dec->canProcess()) { dec->process(); progress = true; }
            }
        }
        delete dec;
        delete demod;
        delete rrc;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
