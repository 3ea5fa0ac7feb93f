// meta.hpp — metadata writer surface of the B200 drop-in modules.
//
// API-compatible with the public part of the reference's include/meta.hpp:10-47 (Serializer, StringSerializer,
// MetaWriter, FileMetaWriter, PipelineMetaWriter), implemented header-only.  The reference's MetaCollector
// (include/meta.hpp:50-66) has no counterpart here: collecting/dirty tracking happens inside libdigiham_b200
// (device events + host replay), the decoders hand finished key/value maps to the MetaWriter.
#pragma once

#include <csdr/source.hpp>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>

namespace Digiham {

    class Serializer {
        public:
            virtual ~Serializer() = default;
            virtual std::string serializeMetaData(std::map<std::string, std::string> metadata) = 0;
    };

    // `key:value;key:value\n`, keys in map order (reference src/lib/meta.cpp:8-17)
    class StringSerializer: public Serializer {
        public:
            std::string serializeMetaData(std::map<std::string, std::string> metadata) override {
                std::string line;
                const char* sep = "";
                for (const auto& kv : metadata) {
                    line += sep;
                    line += kv.first;
                    line += ':';
                    line += kv.second;
                    sep = ";";
                }
                line += '\n';
                return line;
            }
    };

    class MetaWriter {
        public:
            MetaWriter(): MetaWriter(new StringSerializer()) {}
            explicit MetaWriter(Serializer* s): serializer(s) {}
            virtual ~MetaWriter() { delete serializer; }
            virtual void sendMetaData(std::map<std::string, std::string> metadata) = 0;
            void setSerializer(Serializer* s) {
                if (s == serializer) return;
                Serializer* old = serializer;
                serializer = s;
                delete old;
            }
        protected:
            Serializer* serializer;
    };

    // owns and closes the FILE* (reference src/lib/meta.cpp:34-46)
    class FileMetaWriter: public MetaWriter {
        public:
            explicit FileMetaWriter(FILE* out): MetaWriter(), file(out) {}
            FileMetaWriter(FILE* out, Serializer* s): MetaWriter(s), file(out) {}
            ~FileMetaWriter() override { fclose(file); }
            void sendMetaData(std::map<std::string, std::string> metadata) override {
                const std::string text = serializer->serializeMetaData(std::move(metadata));
                fwrite(text.data(), 1, text.size(), file);
                fflush(file);
            }
        private:
            FILE* file = nullptr;
    };

    // writes into a csdr writer, dropping updates that do not fit (reference src/lib/meta.cpp:48-56)
    class PipelineMetaWriter: public MetaWriter, public Csdr::Source<unsigned char> {
        public:
            explicit PipelineMetaWriter(Serializer* s): MetaWriter(s) {}
            void sendMetaData(std::map<std::string, std::string> metadata) override {
                const std::string text = serializer->serializeMetaData(std::move(metadata));
                if (writer == nullptr || writer->writeable() < text.size()) return;
                std::memcpy(writer->getWritePointer(), text.data(), text.size());
                writer->advance(text.size());
            }
    };

}
