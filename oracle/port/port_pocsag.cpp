// port_pocsag.cpp — CPU restatement of the reference's POCSAG decoder.  TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
//
// Follows Digiham::Pocsag::{SyncPhase,CodewordPhase} (reference src/pocsag_decoder/pocsag_phase.cpp:10-92),
// Codeword (codeword.cpp:9-55) and Message (message.cpp:7-73).
#include "port.hpp"

#include <cstring>

namespace port {

namespace {

const uint32_t kFrameSync = 0x7CD215D8u;
const uint32_t kIdleWord = 0x7A89C197u;

bool hasSync(const uint8_t* bits) {
    unsigned distance = 0;
    for (int i = 0; i < 32; i++) {
        const uint8_t expect = (uint8_t) ((kFrameSync >> (31 - i)) & 1u);
        distance += (unsigned) __builtin_popcount((unsigned) (bits[i] ^ expect));
    }
    return distance <= 3;
}

struct Page {
    bool open = false;
    uint32_t address = 0;
    int function = 0;
    char text[80];
    int fill = 0;

    void start(uint32_t a, int f) {
        open = true;
        address = a;
        function = f;
        std::memset(text, 0, sizeof(text));
        fill = 0;
    }
    // Message::append (message.cpp:26-71): only alphanumeric pages (function 3) ever store characters, because
    // pages are only created for functions 1 and 3 (pocsag_phase.cpp:70-75)
    void append(uint32_t payload) {
        if (function != 3) return;
        if (fill + 20 < 80 * 7) {
            for (int i = 0; i < 20; i++) {
                const int bit = (payload >> (19 - i)) & 1;
                text[fill / 7] |= (char) (bit << (fill % 7));
                fill++;
            }
        }
    }
    void flush(Decoded& out) const {
        if (!open || fill == 0) return;
        const std::string line = serialize({{"address", std::to_string(address)}, {"message", std::string(text)}});
        out.bytes.insert(out.bytes.end(), line.begin(), line.end());
    }
};

}  // namespace

void decode_pocsag(const uint8_t* sym, size_t n, Decoded& out) {
    bool inBatch = false;
    int syncCount = 0, wordIndex = 0;
    Page page;
    size_t pos = 0;
    while (n - pos > 32) {
        const uint8_t* p = sym + pos;
        if (!inBatch) {
            if (hasSync(p)) {
                pos += 32;
                inBatch = true;
                syncCount = 1;
                wordIndex = 0;
                page.open = false;
            } else {
                pos++;
            }
            continue;
        }
        if (wordIndex >= 16) {
            if (hasSync(p)) {
                if (syncCount++ > 2) syncCount = 2;
            } else if (syncCount-- < 0) {
                page.flush(out);
                inBatch = false;          // no bits consumed on the way back to the sync search
                continue;
            }
            pos += 32;
            wordIndex = 0;
            continue;
        }
        uint32_t word = 0;
        for (int i = 0; i < 32; i++) word |= (uint32_t) (p[i] && 1) << (31 - i);
        uint32_t upper = word >> 1;
        bool ok = correct(BCH31_21, upper);
        if (ok) {
            word = (word & 1u) | (upper << 1);
            ok = (__builtin_popcount(word) & 1) == 0;      // even parity over all 32 bits
        }
        if (!ok) {
            page.open = false;                            // the page is dropped without being written
        } else if (word == kIdleWord) {
            page.flush(out);
            page.open = false;
        } else if ((word >> 31) == 0) {
            page.flush(out);
            page.open = false;
            const int function = (word >> 11) & 3;
            if (function == 1 || function == 3) page.start((((word >> 13) & 0x3FFFFu) << 3) | (uint32_t) (wordIndex / 2), function);
        } else if (page.open) {
            page.append((word >> 11) & 0xFFFFFu);
        }
        pos += 32;
        wordIndex++;
    }
}

}  // namespace port
