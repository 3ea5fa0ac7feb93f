// digiham_b200_modules.hpp — header-compatible facade classes over libdigiham_b200 (C ABI: digiham_b200.h).
//
// Same namespaces, class names, constructors and Csdr::Module<T,U> surface as the reference's public headers, so a
// pipe graph written against digiham keeps compiling and now runs its hot path on the GPU:
//   Digiham::RrcFilter::{RrcFilter,NarrowRrcFilter,WideRrcFilter}     reference include/rrc_filter.hpp:10-31
//   Digiham::Fsk::FskDemodulator / GfskDemodulator                    reference include/fsk_demodulator.hpp:12-33,
//                                                                     include/gfsk_demodulator.hpp:12-33
//   Digiham::DigitalVoice::DigitalVoiceFilter                         reference include/digitalvoice_filter.hpp:12-19
//   Digiham::Decoder, Dmr::Decoder, Ysf::Decoder, Pocsag::Decoder     reference include/decoder.hpp:17-30,
//                                                                     dmr_decoder.hpp:9-17, ysf_decoder.hpp:9-12,
//                                                                     pocsag_decoder.hpp:9-16
// One module instance is a one-channel bank (BASELINE config 1, "plumbing"); thousands of channels should share one
// bank through the C ABI directly.  There is no CPU fallback: constructors throw
// std::runtime_error when no GPU is available.
//
// Behavioural difference that is NOT observable in the produced streams: where a reference module consumes one
// symbol / one frame per process() call and leaves the rest in the reader, these modules hand everything that is
// buffered to the bank (which carries what it cannot use yet) and keep output that does not fit the writer in a
// pending queue.  `while (m->canProcess()) m->process();` (src/lib/cli.cpp:29-33) produces identical bytes.
#pragma once

#include <csdr/module.hpp>

#include "digiham_b200.h"
#include "meta.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace Digiham {

    namespace B200 {

        inline int deviceIndex() {
            const char* e = std::getenv("DIGIHAM_B200_DEVICE");
            return e ? std::atoi(e) : 0;
        }

        inline void require(int rc, const char* what) {
            if (rc != DH_OK) throw std::runtime_error(std::string(what) + ": " + dh_last_error());
        }

        // bytes produced by a bank that did not fit the csdr writer yet
        class Pending {
            public:
                bool empty() const { return head == data.size(); }
                void append(const unsigned char* p, size_t n) {
                    if (empty()) {
                        data.clear();
                        head = 0;
                    }
                    data.insert(data.end(), p, p + n);
                }
                template <typename W>
                void drainTo(W* writer) {
                    while (!empty()) {
                        const size_t room = writer->writeable();
                        if (room == 0) return;
                        const size_t n = std::min(room, data.size() - head);
                        std::memcpy(writer->getWritePointer(), data.data() + head, n);
                        writer->advance(n);
                        head += n;
                    }
                }
            private:
                std::vector<unsigned char> data;
                size_t head = 0;
        };

    }

    namespace RrcFilter {

        class RrcFilter: public Csdr::AnyLengthModule<float, float> {
            public:
                RrcFilter(unsigned int nZeros, double gain, const float coeffs[]) {
                    B200::require(dh_rrc_create_custom(&bank, B200::deviceIndex(), 1, nZeros, gain, coeffs),
                                  "RrcFilter");
                }
                ~RrcFilter() override {
                    dh_host_scratch_release(bank);
                    dh_rrc_destroy(bank);
                }
                void process(float* input, float* output, size_t length) override {
                    B200::require(dh_rrc_process_host(bank, 1, input, length, output, length, length),
                                  "RrcFilter::process");
                }
            protected:
                explicit RrcFilter(int kind) {
                    B200::require(dh_rrc_create(&bank, B200::deviceIndex(), 1, kind), "RrcFilter");
                }
                size_t maxLength() override { return 1 << 20; }
            private:
                dh_rrc* bank = nullptr;
        };

        class NarrowRrcFilter: public RrcFilter {
            public:
                NarrowRrcFilter(): RrcFilter(DH_RRC_NARROW) {}
        };

        class WideRrcFilter: public RrcFilter {
            public:
                WideRrcFilter(): RrcFilter(DH_RRC_WIDE) {}
        };

    }

    namespace Fsk {

        // shared implementation of the 2- and 4-level demodulators
        class DemodulatorBase: public Csdr::Module<float, unsigned char> {
            public:
                ~DemodulatorBase() override {
                    dh_host_scratch_release(bank);
                    dh_demod_destroy(bank);
                }
                bool canProcess() override {
                    std::lock_guard<std::mutex> lock(this->processMutex);
                    if (this->writer->writeable() == 0) return false;
                    // same threshold as the reference (gfsk_demodulator.cpp:21): one symbol plus the timing "jump"
                    return !pending.empty() || this->reader->available() > samplesPerSymbol + 1;
                }
                void process() override {
                    std::lock_guard<std::mutex> lock(this->processMutex);
                    pending.drainTo(this->writer);
                    if (!pending.empty()) return;
                    const size_t n = std::min<size_t>(this->reader->available(), 1 << 20);
                    if (n == 0) return;
                    symbols.resize(dh_demod_max_symbols(bank, n));
                    uint32_t produced = 0;
                    B200::require(dh_demod_process_host(bank, 1, this->reader->getReadPointer(), n, n, symbols.data(),
                                                        symbols.size(), &produced),
                                  "demodulator process");
                    this->reader->advance(n);
                    pending.append(symbols.data(), produced);
                    pending.drainTo(this->writer);
                }
            protected:
                DemodulatorBase(unsigned int sps, bool fourLevel, bool invert): samplesPerSymbol(sps) {
                    B200::require(dh_demod_create(&bank, B200::deviceIndex(), 1, fourLevel ? 1 : 0, sps, invert ? 1 : 0),
                                  "demodulator");
                }
            private:
                unsigned int samplesPerSymbol;
                dh_demod* bank = nullptr;
                std::vector<unsigned char> symbols;
                B200::Pending pending;
        };

        class FskDemodulator: public DemodulatorBase {
            public:
                explicit FskDemodulator(unsigned int samplesPerSymbol, bool invert = false):
                    DemodulatorBase(samplesPerSymbol, false, invert) {}
        };

        class GfskDemodulator: public DemodulatorBase {
            public:
                explicit GfskDemodulator(unsigned int samplesPerSymbol): DemodulatorBase(samplesPerSymbol, true, false) {}
        };

    }

    namespace DigitalVoice {

        class DigitalVoiceFilter: public Csdr::AnyLengthModule<short, short> {
            public:
                DigitalVoiceFilter() {
                    B200::require(dh_dvf_create(&bank, B200::deviceIndex(), 1), "DigitalVoiceFilter");
                }
                ~DigitalVoiceFilter() override {
                    dh_host_scratch_release(bank);
                    dh_dvf_destroy(bank);
                }
                void process(short* input, short* output, size_t length) override {
                    B200::require(dh_dvf_process_host(bank, 1, input, length, output, length, length),
                                  "DigitalVoiceFilter::process");
                }
            protected:
                size_t maxLength() override { return 1 << 20; }
            private:
                dh_dvf* bank = nullptr;
        };

    }

    class Decoder: public Csdr::Module<unsigned char, unsigned char> {
        public:
            ~Decoder() override {
                dh_host_scratch_release(bank);
                dh_decoder_destroy(bank);
                delete metaWriter;
            }
            bool canProcess() override {
                std::lock_guard<std::mutex> lock(this->processMutex);
                if (!pending.empty()) return this->writer->writeable() > 0;
                return this->reader->available() > 0;
            }
            void process() override {
                std::lock_guard<std::mutex> lock(this->processMutex);
                pending.drainTo(this->writer);
                if (!pending.empty()) return;
                const uint32_t n = (uint32_t) std::min<size_t>(this->reader->available(), 1 << 20);
                if (n == 0) return;
                B200::require(dh_decoder_process_host(bank, 1, this->reader->getReadPointer(), n, &n), "decoder process");
                this->reader->advance(n);
                const uint8_t* data = nullptr;
                size_t len = 0;
                B200::require(dh_decoder_output(bank, 0, &data, &len), "decoder output");
                if (len) handleOutput(data, len);
                B200::require(dh_decoder_meta_kv(bank, 0, &data, &len), "decoder meta");
                if (len && metaWriter) forwardMeta(data, len);
                dh_decoder_clear(bank);
                pending.drainTo(this->writer);
            }
            // takes ownership, like the reference (src/lib/decoder.cpp:34-40)
            void setMetaWriter(MetaWriter* meta) {
                if (!hasMetaPlane) {
                    delete meta;
                    return;
                }
                delete metaWriter;
                metaWriter = meta;
            }
        protected:
            Decoder(int proto, bool hasMetaPlane): hasMetaPlane(hasMetaPlane) {
                B200::require(dh_decoder_create(&bank, B200::deviceIndex(), 1, proto), "Decoder");
            }
            virtual void handleOutput(const uint8_t* data, size_t len) { pending.append(data, len); }
            dh_decoder* bank = nullptr;
            B200::Pending pending;
        private:
            void forwardMeta(const uint8_t* p, size_t len) {
                size_t pos = 0;
                auto get16 = [&]() -> size_t {
                    const size_t v = p[pos] | (size_t) p[pos + 1] << 8;
                    pos += 2;
                    return v;
                };
                while (pos + 2 <= len) {
                    std::map<std::string, std::string> update;
                    const size_t pairs = get16();
                    for (size_t i = 0; i < pairs; i++) {
                        const size_t kl = get16();
                        std::string key((const char*) p + pos, kl);
                        pos += kl;
                        const size_t vl = get16();
                        update[key] = std::string((const char*) p + pos, vl);
                        pos += vl;
                    }
                    metaWriter->sendMetaData(update);
                }
            }
            bool hasMetaPlane;
            MetaWriter* metaWriter = nullptr;
    };

    namespace Dmr {

        class Decoder: public Digiham::Decoder {
            public:
                Decoder(): Digiham::Decoder(DH_PROTO_DMR, true) {}
                // bit 0 = slot 1 audible, bit 1 = slot 2 audible (reference include/dmr_decoder.hpp:12)
                void setSlotFilter(unsigned char filter) { dh_decoder_set_slot_filter(bank, -1, filter); }
        };

    }

    namespace Ysf {

        class Decoder: public Digiham::Decoder {
            public:
                Decoder(): Digiham::Decoder(DH_PROTO_YSF, true) {}
        };

    }

    namespace Nxdn {

        // reference include/nxdn_decoder.hpp:9-14
        class Decoder: public Digiham::Decoder {
            public:
                Decoder(): Digiham::Decoder(DH_PROTO_NXDN, true) {}
        };

    }

    namespace DStar {

        // reference include/dstar_decoder.hpp:9-13
        class Decoder: public Digiham::Decoder {
            public:
                Decoder(): Digiham::Decoder(DH_PROTO_DSTAR, true) {}
        };

    }

    namespace Pocsag {

        class Decoder: public Digiham::Decoder {
            public:
                Decoder(): Decoder(new StringSerializer()) {}
                // like the reference, the serializer is never freed (src/pocsag_decoder/pocsag_decoder.cpp:6-8)
                explicit Decoder(Serializer* serializer): Digiham::Decoder(DH_PROTO_POCSAG, false), serializer(serializer) {}
            protected:
                // The bank renders `address:N;message:TEXT\n` (StringSerializer) on the device.  For another
                // serializer the rendered text is dropped and the structured {address, message} records the bank
                // keeps beside it (dh_decoder_meta_kv) are serialised instead — no re-parsing of the text, so message
                // bodies with ';', ':' or newlines survive (reference src/pocsag_decoder/message.cpp:16-24).
                void handleOutput(const uint8_t* data, size_t len) override {
                    if (dynamic_cast<StringSerializer*>(serializer) != nullptr) {
                        pending.append(data, len);
                        return;
                    }
                    const uint8_t* p = nullptr;
                    size_t n = 0;
                    B200::require(dh_decoder_meta_kv(bank, 0, &p, &n), "pocsag records");
                    size_t pos = 0;
                    auto get16 = [&]() -> size_t {
                        const size_t v = p[pos] | (size_t) p[pos + 1] << 8;
                        pos += 2;
                        return v;
                    };
                    while (pos + 2 <= n) {
                        std::map<std::string, std::string> record;
                        const size_t pairs = get16();
                        for (size_t i = 0; i < pairs; i++) {
                            const size_t kl = get16();
                            std::string key((const char*) p + pos, kl);
                            pos += kl;
                            const size_t vl = get16();
                            record[key] = std::string((const char*) p + pos, vl);
                            pos += vl;
                        }
                        const std::string out = serializer->serializeMetaData(record);
                        pending.append((const unsigned char*) out.data(), out.size());
                    }
                }
                Serializer* serializer;
        };

    }

}
