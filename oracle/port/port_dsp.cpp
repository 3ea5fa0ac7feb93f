// port_dsp.cpp — CPU restatement ("port") of the sample-rate DSP modules of the reference.
//
// TEST INFRASTRUCTURE ONLY: this is the checker, never the product.  Nothing under digiham_b200/ or include/
// may include, link or call it.  It exists so that parity can be checked on machines where the compiled
// reference (oracle/_ref) is not available; it is itself pinned against the compiled reference and against
// tests/golden/ (tests/test_oracle_cpu.py).
//
// Every function follows the reference's arithmetic literally (same operation order, same float/double
// promotions); build flags mirror the reference build (-O3, x86-64 baseline, -ffp-contract=off: no FMA).
#include "port.hpp"
#include "rrc_taps.inc"

#include <cfloat>
#include <cmath>

namespace port {

// ---- RrcFilter (reference src/rrc_filter/rrc_filter.cpp:16-34) ----------------------------------------------------
Rrc::Rrc(bool narrow):
    zeros(narrow ? 160 : 80),
    gain(narrow ? 1.667711971e+01 : 8.337797030e+00),
    taps(narrow ? kNarrowTaps : kWideTaps),
    line(zeros + 1, 0.0f)   // the reference leaves the delay line uninitialised (:9); the contract says zeros
{}

float Rrc::step(float sample) {
    // shift register, newest sample last (:25-28)
    for (unsigned i = 0; i < zeros; i++) line[i] = line[i + 1];
    line[zeros] = sample;
    // strictly ordered float accumulation of separately rounded products (:30-31)
    float acc = 0.0f;
    for (unsigned i = 0; i <= zeros; i++) acc += taps[i] * line[i];
    // float / double -> double division, narrowed on return (:33)
    return (float) (acc / gain);
}

// ---- GfskDemodulator / FskDemodulator (reference src/gfsk_demodulator/gfsk_demodulator.cpp:18-122,
//      src/fsk_demodulator/fsk_demodulator.cpp:19-112) --------------------------------------------------------------
Demod::Demod(unsigned sps, bool fourLevel, bool invert):
    sps(sps), fourLevel(fourLevel), invert(invert),
    evalFrom((unsigned) (int) roundf((float) sps / 3)),
    evalTo((unsigned) (int) roundf((float) sps * 2 / 3)),
    history(100 * sps, 0.0f),
    volumes(100, 0.0f)     // volume_rb is uninitialised in the reference (include/gfsk_demodulator.hpp:27)
{}

void Demod::run(const float* x, size_t n, std::vector<uint8_t>& out) {
    size_t rd = 0;
    // canProcess(): more than sps + 1 samples buffered (:21)
    while (n - rd > sps + 1) {
        const float* in = x + rd;
        float sum = 0.0f, volumeSum = 0.0f;
        for (unsigned i = 0; i < sps; i++) {
            const float v = in[i];
            if (i >= evalFrom && i < evalTo) sum += v;
            volumeSum += v;
            history[historyPos + i] = v;
        }
        rd += sps + nudge;       // advance(samplesPerSymbol + variance_offset) (:36)
        nudge = 0;
        historyPos += sps;
        if (historyPos >= history.size()) {
            // every 100 symbols: sample phase with the smallest variance (:41-66)
            double best = 0;
            size_t bestPhase = 0;
            for (unsigned ph = 0; ph < sps; ph++) {
                float total = 0;
                for (int k = 0; k < 100; k++) total += history[k * sps + ph];
                const double mean = total / 100;            // float division, then widened (:53)
                double dev = 0;
                for (int k = 0; k < 100; k++) {
                    const double d = mean - history[k * sps + ph];
                    dev += d * d;                           // pow(x, 2) (:58)
                }
                const double variance = dev / 100;
                if (ph == 0 || variance < best) {
                    best = variance;
                    bestPhase = ph;
                }
            }
            if (best <= 0 || best > 5000000) {
                // no decision (:69-70)
            } else if (bestPhase > 0 && bestPhase < sps / 2) {
                nudge = +1;
            } else if (bestPhase >= sps / 2 && bestPhase < sps - 1) {
                nudge = -1;
            }
            historyPos %= history.size();
        }
        volumes[volumePos] = volumeSum / sps;
        if (++volumePos >= volumes.size()) volumePos = 0;
        // calibrateAudio() (:109-122): note FLT_MIN, the smallest POSITIVE float, as the initial maximum
        float lo = FLT_MAX, hi = FLT_MIN;
        for (float v : volumes) {
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
        const float center = (hi + lo) / 2;
        const float average = sum / (evalTo - evalFrom);
        uint8_t symbol;
        if (fourLevel) {
            const float upper = (hi - center) * 0.625 + center;   // double arithmetic, narrowed (:120-121)
            const float lower = (lo - center) * 0.625 + center;
            if (average > center) symbol = average > upper ? 1 : 0;
            else symbol = average < lower ? 3 : 2;
        } else {
            symbol = average > center ? !invert : invert;         // fsk_demodulator.cpp:93-97
        }
        out.push_back(symbol);
    }
}

// ---- DigitalVoiceFilter (reference src/digitalvoice_filter/digitalvoice_filter.cpp:6-45) ---------------------------
short Dvf::step(short in) {
    const float sample = (float) in / 32767;     // SHRT_MAX
    for (int i = 0; i < 10; i++) {
        xv[i] = xv[i + 1];
        yv[i] = yv[i + 1];
    }
    xv[10] = sample / 5;                         // GAIN, hand-tuned (:29-32)
    yv[10] = (xv[10] - xv[0]) + 5 * (xv[2] - xv[8]) + 10 * (xv[6] - xv[4])
             + (0.1254306222 * yv[0]) + (0.1285714097 * yv[1])
             + (-0.8106454980 * yv[2]) + (-0.7664515771 * yv[3])
             + (2.1846187758 * yv[4]) + (1.8106678608 * yv[5])
             + (-3.1465011600 * yv[6]) + (-2.0391991609 * yv[7])
             + (2.4873968618 * yv[8]) + (1.0249072542 * yv[9]);
    return (short) (yv[10] * 32767);
}

}  // namespace port
