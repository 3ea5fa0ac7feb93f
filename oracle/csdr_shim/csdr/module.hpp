// csdr shim — TEST INFRASTRUCTURE ONLY (see reader.hpp for provenance).
//
// Module<T,U>: canProcess()/process() virtuals, protected reader/writer/processMutex
//   (include/decoder.hpp:22-23, src/lib/decoder.cpp:22-28).
// AnyLengthModule<T,U>: process() = min(available, writeable) items through the 3-argument hook
//   (include/rrc_filter.hpp:14, include/digitalvoice_filter.hpp:14); SURVEY.md appendix A.5.
#pragma once
#include "sink.hpp"
#include "source.hpp"
#include <mutex>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cmath>    // the reference relies on <cmath> arriving through this header (gfsk_demodulator.cpp:8,58)

namespace Csdr {

    template <typename T, typename U>
    class Module: public Sink<T>, public Source<U> {
        public:
            ~Module() override = default;
            virtual bool canProcess() = 0;
            virtual void process() = 0;
        protected:
            std::mutex processMutex;
    };

    template <typename T, typename U>
    class AnyLengthModule: public Module<T, U> {
        public:
            bool canProcess() override {
                std::lock_guard<std::mutex> lock(this->processMutex);
                return workSize() > 0;
            }
            void process() override {
                std::lock_guard<std::mutex> lock(this->processMutex);
                size_t n = workSize();
                process(this->reader->getReadPointer(), this->writer->getWritePointer(), n);
                this->reader->advance(n);
                this->writer->advance(n);
            }
        protected:
            virtual void process(T* input, U* output, size_t length) = 0;
            virtual size_t maxLength() { return SIZE_MAX; }
        private:
            size_t workSize() {
                return std::min({this->reader->available(), this->writer->writeable(), maxLength()});
            }
    };

}
