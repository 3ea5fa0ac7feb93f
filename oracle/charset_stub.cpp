// charset_stub.cpp — stands in for src/lib/charset.cpp (ICU, not installed) in the oracle build.
// TEST INFRASTRUCTURE ONLY.  The only conversion the hot path ever asks for is the default
// iso-8859-1 -> utf-8 (src/lib/charset.hpp:11), which is a fixed 1- or 2-byte expansion; like the ICU
// path (result = std::string(target), src/lib/charset.cpp:24) the result stops at the first NUL.
#include "charset.hpp"

using namespace Digiham;

std::string Converter::convertToUtf8(const char* input, const size_t length, const char* /*charset*/) {
    std::string result;
    for (size_t i = 0; i < length; i++) {
        unsigned char c = (unsigned char) input[i];
        if (c == 0) break;
        if (c < 0x80) {
            result.push_back((char) c);
        } else {
            result.push_back((char) (0xC0 | (c >> 6)));
            result.push_back((char) (0x80 | (c & 0x3F)));
        }
    }
    return result;
}
