"""Seeded synthetic signal generators (harness glue for tests/ and bench.py — not part of the product path).

Signals follow SURVEY.md §7 step 2 / §8d: 4-level (or 2-level) NRZ with rectangular hold at `sps` samples per
symbol, float32 in [-1, 1], the format `rrc_filter` receives from the FM discriminator
(reference examples/dmr-decoder.sh:13-19).  Optional impairments: AWGN, DC offset, sampling-clock offset (ppm)
and an arbitrary start phase, so that the timing recovery of the demodulator (variance_offset = +-1) fires.
"""
import numpy as np

# dibit -> deviation level, reference src/gfsk_demodulator/gfsk_demodulator.cpp:92-104 (01 -> +3, 00 -> +1,
# 10 -> -1, 11 -> -3)
LEVELS4 = np.array([1.0, 3.0, -1.0, -3.0], dtype=np.float64) / 3.0
# bit -> level for the 2-level demodulator; with invert=True a 1 is the LOW level (examples/pocsag-decoder.sh:20)
LEVELS2 = np.array([-1.0, 1.0], dtype=np.float64)


def modulate(symbols, sps=10, n_samples=None, amplitude=0.5, levels=LEVELS4, ppm=0.0, phase=0.0, snr_db=None,
             dc=0.0, rng=None):
    """symbols: 1-D int array.  Returns float32 samples (rectangular hold).

    Sample t shows symbol floor((t + phase) * (1 + ppm * 1e-6) / sps); samples beyond the last symbol repeat it.
    """
    symbols = np.asarray(symbols)
    if n_samples is None:
        n_samples = symbols.size * sps
    t = np.arange(n_samples, dtype=np.float64)
    idx = np.floor((t + phase) * (1.0 + ppm * 1e-6) / sps).astype(np.int64)
    idx = np.clip(idx, 0, symbols.size - 1)
    x = levels[symbols[idx]] * amplitude + dc
    if snr_db is not None:
        if rng is None:
            rng = np.random.default_rng(0)
        # signal power of equiprobable levels
        p_sig = float(np.mean(levels ** 2)) * amplitude * amplitude
        sigma = np.sqrt(p_sig / (10.0 ** (snr_db / 10.0)))
        x = x + rng.normal(0.0, sigma, size=n_samples)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def random_symbols(n, levels=4, seed=0):
    return np.random.default_rng(seed).integers(0, levels, size=n).astype(np.uint8)
