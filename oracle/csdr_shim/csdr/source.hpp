// csdr shim — TEST INFRASTRUCTURE ONLY (see reader.hpp for provenance).
#pragma once
#include "writer.hpp"

namespace Csdr {

    template <typename T>
    class Source {
        public:
            virtual ~Source() = default;
            // call sites: src/lib/cli.cpp:27, include/meta.hpp:42 (PipelineMetaWriter uses the `writer` member)
            virtual void setWriter(Writer<T>* w) { writer = w; }
            virtual Writer<T>* getWriter() { return writer; }
            virtual bool hasWriter() { return writer != nullptr; }
        protected:
            Writer<T>* writer = nullptr;
    };

}
