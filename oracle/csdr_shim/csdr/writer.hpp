// csdr shim — TEST INFRASTRUCTURE ONLY (see reader.hpp for provenance).
#pragma once
#include <cstddef>

namespace Csdr {

    template <typename T>
    class Writer {
        public:
            virtual ~Writer() = default;
            virtual size_t writeable() = 0;
            virtual T* getWritePointer() = 0;
            virtual void advance(size_t how_much) = 0;
    };

}
