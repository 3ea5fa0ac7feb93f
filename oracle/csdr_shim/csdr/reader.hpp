// csdr shim — TEST INFRASTRUCTURE ONLY (oracle build), not product code.
//
// libcsdr (module "Csdr" >= 0.18, /root/reference/CMakeLists.txt:17) is neither vendored in the
// reference tree nor installed in this image.  These headers reconstruct the minimal dataflow
// contract the reference hot path relies on, from its call sites only:
//   Reader<T>::available/getReadPointer/advance   (src/gfsk_demodulator/gfsk_demodulator.cpp:21,26,36)
//   Writer<T>::writeable/getWritePointer/advance  (src/gfsk_demodulator/gfsk_demodulator.cpp:21,94,106)
// ABI compatibility with a real libcsdr.so is NOT claimed.
#pragma once
#include <cstddef>

namespace Csdr {

    template <typename T>
    class Reader {
        public:
            virtual ~Reader() = default;
            virtual size_t available() = 0;
            // must expose at least available() contiguous items
            virtual T* getReadPointer() = 0;
            virtual void advance(size_t how_much) = 0;
            virtual void wait() {}
            virtual void unblock() {}
    };

}
