// meta.hpp — metadata writer surface of the B200 drop-in modules.
//
// API-compatible with the reference's include/meta.hpp:10-66 (Serializer, StringSerializer, MetaWriter,
// FileMetaWriter, PipelineMetaWriter, MetaCollector), implemented header-only.  The B200 decoders do their own
// collecting / dirty tracking inside libdigiham_b200 (device events + host replay) and hand finished key/value maps
// to the MetaWriter; MetaCollector is kept for code that derives its own collectors from it.
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

    // Base of the per-protocol collectors (reference include/meta.hpp:50-66, src/lib/meta.cpp:58-100): subclasses
    // call sendMetaData() whenever a field changed; between hold() and the matching release() updates are coalesced
    // into one (the collector is only marked dirty) and sent when the last hold is released.  Owns its writer.
    class MetaCollector {
        public:
            MetaCollector(): MetaCollector(nullptr) {}
            explicit MetaCollector(MetaWriter* w): writer(w) {}
            virtual ~MetaCollector() { delete writer; }
            void setWriter(MetaWriter* w) {
                delete writer;
                writer = w;
            }
            void hold() { held++; }
            void release() {
                if (--held != 0) return;
                const bool pending = dirty;
                dirty = false;
                if (pending) sendMetaData();
            }
        protected:
            virtual std::string getProtocol() = 0;
            virtual std::map<std::string, std::string> collect() {
                std::map<std::string, std::string> fields;
                fields["protocol"] = getProtocol();
                return fields;
            }
            virtual void sendMetaData() {
                if (writer == nullptr) return;
                if (held) dirty = true;
                else sendMetaData(collect());
            }
            virtual void sendMetaData(std::map<std::string, std::string> metadata) {
                if (writer != nullptr) writer->sendMetaData(std::move(metadata));
            }
        private:
            int held = 0;
            bool dirty = false;
            MetaWriter* writer;
    };

}
