// port.hpp — declarations of the CPU restatement ("port").  TEST INFRASTRUCTURE ONLY (see port_dsp.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace port {

// ---- sample-rate DSP (port_dsp.cpp)
class Rrc {
    public:
        explicit Rrc(bool narrow);
        float step(float sample);
    private:
        unsigned zeros;
        double gain;
        const float* taps;
        std::vector<float> line;
};

class Demod {
    public:
        Demod(unsigned sps, bool fourLevel, bool invert);
        // consumes as much of x[0..n) as the reference would with everything buffered; appends symbols
        void run(const float* x, size_t n, std::vector<uint8_t>& out);
    private:
        unsigned sps;
        bool fourLevel, invert;
        unsigned evalFrom, evalTo;
        std::vector<float> history;
        size_t historyPos = 0;
        int nudge = 0;
        std::vector<float> volumes;
        size_t volumePos = 0;
};

class Dvf {
    public:
        short step(short in);
    private:
        float xv[11] = {0};
        float yv[11] = {0};
};

// ---- block codes and friends (port_fec.cpp)
enum CodeId { H7_4 = 0, H13_9, H15_11, H16_11, QR16_7, GOLAY20_8, GOLAY24_12, BCH31_21, kNumCodes };
uint32_t syndrome(int code, uint32_t word);
bool correct(int code, uint32_t& word);            // false = uncorrectable, word untouched
bool bptc_196_96(const uint8_t payload[25], uint8_t out[12]);
unsigned viterbi(const uint8_t* packedDibits, unsigned steps, uint8_t* out);   // returns the best metric (mod 256)
uint16_t crc16(const uint8_t* data, int count);
void dewhiten(const uint8_t* in, uint8_t* out, unsigned nbits);
unsigned hamming_distance(const uint8_t* a, const uint8_t* b, size_t n);

std::string serialize(const std::map<std::string, std::string>& kv);            // StringSerializer
std::string latin1_to_utf8(const unsigned char* p, size_t n);

// ---- protocol decoders: whole symbol stream in, byte stream + metadata lines out
struct Decoded {
    std::vector<uint8_t> bytes;
    std::string meta;
};
void decode_dmr(const uint8_t* sym, size_t n, int slotFilter, Decoded& out);
void decode_ysf(const uint8_t* sym, size_t n, Decoded& out);
void decode_pocsag(const uint8_t* sym, size_t n, Decoded& out);
void decode_nxdn(const uint8_t* sym, size_t n, Decoded& out);
void decode_dstar(const uint8_t* sym, size_t n, Decoded& out);
int dstar_header_probe(const uint8_t* raw660, char* text, size_t cap);
unsigned nxdn_viterbi(const uint8_t* packedDibits, unsigned nbits, uint8_t* out);
int nxdn_sacch_probe(const uint8_t* dibits30, uint8_t out5[5]);
int nxdn_facch1_probe(const uint8_t* dibits72);

}  // namespace port
