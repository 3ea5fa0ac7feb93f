// csdr shim — TEST INFRASTRUCTURE ONLY (see reader.hpp for provenance).
//
// Ringbuffer<T>/RingbufferReader<T>/StdoutWriter<U> are only needed by the reference CLI template
// (src/lib/cli.cpp:10,26-27,101-106).  Upstream maps the ring twice so that readers always see a
// contiguous window; this shim gets the same guarantee with a mirrored second half.
#pragma once
#include "reader.hpp"
#include "writer.hpp"
#include <vector>
#include <cstdio>
#include <cstring>

namespace Csdr {

    template <typename T> class RingbufferReader;

    template <typename T>
    class Ringbuffer: public Writer<T> {
        public:
            explicit Ringbuffer(size_t size): size(size), data(2 * size) {}
            size_t writeable() override { return size - 1 - (written - minRead()); }
            T* getWritePointer() override { return data.data() + (written % size); }
            void advance(size_t n) override {
                // the producer wrote n items starting at the physical index written % size, possibly
                // running into the mirror half; make both halves coherent again
                size_t start = written % size;
                for (size_t i = 0; i < n; i++) {
                    size_t phys = start + i;
                    if (phys >= size) data[phys - size] = data[phys];
                    else data[phys + size] = data[phys];
                }
                written += n;
            }
        private:
            friend class RingbufferReader<T>;
            size_t minRead() { return readPos; }
            size_t size;
            std::vector<T> data;
            size_t written = 0;
            size_t readPos = 0;
    };

    template <typename T>
    class RingbufferReader: public Reader<T> {
        public:
            explicit RingbufferReader(Ringbuffer<T>* rb): rb(rb) {}
            size_t available() override { return rb->written - rb->readPos; }
            T* getReadPointer() override { return rb->data.data() + (rb->readPos % rb->size); }
            void advance(size_t n) override { rb->readPos += n; }
        private:
            Ringbuffer<T>* rb;
    };

    template <typename T>
    class StdoutWriter: public Writer<T> {
        public:
            StdoutWriter(): buffer(1 << 16) {}
            size_t writeable() override { return buffer.size(); }
            T* getWritePointer() override { return buffer.data(); }
            void advance(size_t n) override { fwrite(buffer.data(), sizeof(T), n, stdout); fflush(stdout); }
        private:
            std::vector<T> buffer;
    };

}
