// meta_collector.cpp — behaviour of Digiham::MetaCollector (include/meta.hpp) as the reference defines it
// (reference include/meta.hpp:50-66, src/lib/meta.cpp:58-100): protocol field, writer ownership, hold / release
// coalescing.  Host-only; prints the lines a FileMetaWriter wrote, the test compares them with the expected text.
#include "meta.hpp"

#include <cstdio>
#include <string>

namespace {

class Probe: public Digiham::MetaCollector {
    public:
        using Digiham::MetaCollector::MetaCollector;
        void setCall(const std::string& c) {
            call = c;
            sendMetaData();
        }
        void raw() { Digiham::MetaCollector::sendMetaData({{"x", "1"}}); }
    protected:
        std::string getProtocol() override { return "PROBE"; }
        std::map<std::string, std::string> collect() override {
            auto m = Digiham::MetaCollector::collect();
            if (!call.empty()) m["call"] = call;
            return m;
        }
    private:
        std::string call;
};

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    Probe p;                         // no writer: every update is dropped
    p.setCall("NOBODY");
    p.setWriter(new Digiham::FileMetaWriter(fopen(argv[1], "w")));
    p.setCall("DL1ABC");             // sent at once
    p.hold();
    p.hold();
    p.setCall("DL2DEF");             // held: only marks the collector dirty
    p.setCall("DL3GHI");
    p.release();                     // one hold left: still nothing
    p.raw();                         // the map overload bypasses hold
    p.release();                     // last hold released: ONE update with the latest state
    p.hold();
    p.release();                     // nothing pending: nothing is sent
    p.setCall("DL4JKL");
    return 0;                        // the collector deletes the writer, which closes the file
}
