// csdr shim — TEST INFRASTRUCTURE ONLY (see reader.hpp for provenance).
#pragma once
#include "reader.hpp"

namespace Csdr {

    template <typename T>
    class Sink {
        public:
            virtual ~Sink() = default;
            // call site: src/lib/cli.cpp:26
            virtual void setReader(Reader<T>* r) { reader = r; }
            virtual Reader<T>* getReader() { return reader; }
            virtual bool hasReader() { return reader != nullptr; }
        protected:
            Reader<T>* reader = nullptr;
    };

}
