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


# ---------------------------------------------------------------------------------------------------------------
# Systematic block-code encoders.  For every code of the reference the check matrix is H = [A | I_r]
# (identity in the low r bits), hence codeword = (data << r) | syndrome(data << r)  (SURVEY.md appendix A.0).
_P = {
    "hamming_7_4": (7, 4, ["101", "111", "110", "011"]),
    "hamming_13_9": (13, 9, ["1111", "1110", "0111", "1010", "0101", "1011", "1100", "0110", "0011"]),
    "hamming_15_11": (15, 11, ["1001", "1101", "1111", "1110", "0111", "1010", "0101", "1011", "1100", "0110", "0011"]),
    "hamming_16_11": (16, 11, ["10011", "11010", "11111", "11100", "01110", "10101", "01011", "10110", "11001",
                               "01101", "00111"]),
    "qr_16_7": (16, 7, ["001001111", "100011110", "110110111", "111100010", "111001001", "011100101", "001110011"]),
    "golay_20_8": (20, 8, ["001111011010", "110110011001", "011011001101", "001101100111", "110111000110",
                           "101010010111", "100100111110", "100011101011"]),
    "golay_24_12": (24, 12, ["110001110101", "011000111011", "111101101000", "011110110100", "001111011010",
                             "110110011001", "011011001101", "001101100111", "110111000110", "101010010111",
                             "100100111110", "100011101011"]),
}


def encode_block(code, data):
    """data: k-bit integer (MSB = first data bit) -> n-bit systematic codeword."""
    n, k, rows = _P[code]
    r = n - k
    parity = 0
    for d in range(k):
        if (data >> (k - 1 - d)) & 1:
            parity ^= int(rows[d], 2)
    return (data << r) | parity


def bch_31_21_encode(data21):
    """POCSAG BCH(31,21), generator x^10+x^9+x^8+x^6+x^5+x^3+1."""
    poly = 0b11101101001
    v = data21 << 10
    for s in range(30, 9, -1):
        if v & (1 << s):
            v ^= poly << (s - 10)
    return (data21 << 10) | v


def _bits_to_dibits(bits):
    bits = np.asarray(bits, dtype=np.uint8)
    return (bits[0::2] << 1 | bits[1::2]).astype(np.uint8)


def _int_to_bits(v, n):
    return np.array([(v >> (n - 1 - i)) & 1 for i in range(n)], dtype=np.uint8)


def _hex_to_dibits(h):
    v = int(h, 16)
    n = len(h) * 4
    return _bits_to_dibits(_int_to_bits(v, n))


# ---------------------------------------------------------------------------------------------------------------
# DMR (ETSI TS 102 361-1) base-station frame generator.  Layout as the reference reads it
# (src/dmr_decoder/dmr_phase.cpp:65-302): 144 dibits =
#   CACH(12) | payload(54) | SYNC-or-EMB(24) | payload(54)
DMR_SYNC = {
    "bs_data": _hex_to_dibits("DFF57D75DF5D"),
    "bs_voice": _hex_to_dibits("755FD7DF75F7"),
    "ms_data": _hex_to_dibits("D5D7F77FD757"),
    "ms_voice": _hex_to_dibits("7F7D5DD57DFD"),
}
DMR_DT_VOICE_LC, DMR_DT_TERMINATOR_LC, DMR_DT_CSBK, DMR_DT_RATE_34, DMR_DT_IDLE = 1, 2, 3, 8, 9


def dmr_cach(slot, lcss, rng, busy=1):
    tact4 = (busy << 3) | (slot << 2) | lcss
    tact = encode_block("hamming_7_4", tact4)
    bits = rng.integers(0, 2, size=24).astype(np.uint8)
    for i, pos in enumerate([0, 4, 8, 12, 14, 18, 22]):
        bits[pos] = (tact >> (6 - i)) & 1
    return _bits_to_dibits(bits)


def dmr_bptc_encode(info12):
    """12 info bytes -> 196 transmitted bits (BPTC(196,96), reference src/dmr_decoder/bptc_196_96.c:12-56)."""
    info_bits = np.unpackbits(np.asarray(info12, dtype=np.uint8))
    m = np.zeros((13, 15), dtype=np.uint8)
    m[0, 3:11] = info_bits[:8]
    m[1:9, 0:11] = info_bits[8:].reshape(8, 11)
    for r in range(9):
        data = int("".join(str(b) for b in m[r, :11]), 2)
        cw = encode_block("hamming_15_11", data)
        m[r, :] = _int_to_bits(cw, 15)
    for c in range(15):
        data = int("".join(str(b) for b in m[:9, c]), 2)
        cw = encode_block("hamming_13_9", data)
        m[:, c] = _int_to_bits(cw, 13)
    d = np.zeros(196, dtype=np.uint8)
    d[1:] = m.reshape(-1)
    t = np.zeros(196, dtype=np.uint8)
    for i in range(196):
        t[(i * 181) % 196] = d[i]
    return t


def dmr_slot_type(color_code, data_type):
    return _int_to_bits(encode_block("golay_20_8", (color_code << 4) | data_type), 20)


def dmr_data_burst(slot, data_type, info12, rng, color_code=1, sync="bs_data", lcss=0):
    t = dmr_bptc_encode(info12)
    st = dmr_slot_type(color_code, data_type)
    f = np.zeros(144, dtype=np.uint8)
    f[0:12] = dmr_cach(slot, lcss, rng)
    f[12:61] = _bits_to_dibits(t[:98])
    f[61:66] = _bits_to_dibits(st[:10])
    f[66:90] = DMR_SYNC[sync]
    f[90:95] = _bits_to_dibits(st[10:])
    f[95:144] = _bits_to_dibits(t[98:])
    return f


def dmr_embedded_lc_fragments(lc9):
    """9 LC bytes -> four 32-bit fragments (16 dibits each), reference src/dmr_decoder/embedded.cpp:37-89."""
    lc9 = [int(b) for b in lc9]
    bits = np.unpackbits(np.asarray(lc9, dtype=np.uint8))   # 72 bits
    cs = sum(lc9) % 31
    csbits = _int_to_bits(cs, 5)
    rows = []
    pos = 0
    for r in range(7):
        if r < 2:
            data = bits[pos:pos + 11]
            pos += 11
        else:
            data = np.concatenate([bits[pos:pos + 10], csbits[r - 2:r - 1]])
            pos += 10
        rows.append(encode_block("hamming_16_11", int("".join(str(b) for b in data), 2)))
    parity = 0
    for r in rows:
        parity ^= r
    rows.append(parity)
    data = np.zeros(16, dtype=np.uint8)
    for i in range(16):
        b = 0
        for k in range(8):
            b |= ((rows[k] >> (15 - i)) & 1) << (7 - k)
        data[i] = b
    return [_bits_to_dibits(np.unpackbits(data[4 * q:4 * q + 4])) for q in range(4)]


def dmr_voice_burst(slot, rng, sync="bs_voice", emb=None, fragment=None, lcss_cach=0):
    """emb = (color_code, pi, lcss) for superframe bursts B..F, None for burst A (voice sync)."""
    f = np.zeros(144, dtype=np.uint8)
    f[0:12] = dmr_cach(slot, lcss_cach, rng)
    f[12:66] = rng.integers(0, 4, size=54)
    f[90:144] = rng.integers(0, 4, size=54)
    if emb is None:
        f[66:90] = DMR_SYNC[sync]
    else:
        cc, pi, lcss = emb
        e = _int_to_bits(encode_block("qr_16_7", (cc << 3) | (pi << 2) | lcss), 16)
        f[66:70] = _bits_to_dibits(e[:8])
        f[70:86] = fragment if fragment is not None else rng.integers(0, 4, size=16)
        f[86:90] = _bits_to_dibits(e[8:])
    return f


def _gf256_mul(a, b):
    """GF(2^8) with the field polynomial x^8 + x^4 + x^3 + x^2 + 1 (ETSI TS 102 361-1 B.3.6)."""
    acc = 0
    for i in range(8):
        if (b >> i) & 1:
            acc ^= a << i
    for i in range(14, 7, -1):
        if acc & (1 << i):
            acc ^= 0x11D << (i - 8)
    return acc


def dmr_rs_12_9_parity(lc9, mask=0x96):
    """The three Reed-Solomon (12,9) parity octets of a 9-octet full LC, generator (x + a)(x + a^2)(x + a^3) =
    x^3 + 14 x^2 + 56 x + 64, XOR-ed with the data-type mask (0x96 voice LC header, 0x99 terminator with LC)."""
    reg = [0, 0, 0]                      # remainder, highest degree first
    for d in lc9:
        fb = d ^ reg[0]
        reg = [reg[1] ^ _gf256_mul(fb, 14), reg[2] ^ _gf256_mul(fb, 56), _gf256_mul(fb, 64)]
    return [r ^ mask for r in reg]


def dmr_full_lc(opcode, target, source, fid=0, options=0, rng=None):
    """9 LC octets + 3 tail octets.  The reference never looks at the tail (src/dmr_decoder/lc.cpp:8-11), so the
    default generator fills it with random octets; dmr_rs_12_9_parity makes a standard-conforming tail."""
    lc = [opcode & 0x3F, fid, options, (target >> 16) & 0xFF, (target >> 8) & 0xFF, target & 0xFF,
          (source >> 16) & 0xFF, (source >> 8) & 0xFF, source & 0xFF]
    tail = list(rng.integers(0, 256, size=3)) if rng is not None else [0, 0, 0]
    return lc + [int(t) for t in tail]


def dmr_slot_script(rng, kind):
    """A list of burst makers for one TDMA slot.  kind: 'voice', 'data', 'idle', 'mixed'."""
    bursts = []

    def voice_call(n_super, target, source, group=True, with_ta=False, with_gps=False):
        opcode = 0 if group else 3
        lc = dmr_full_lc(opcode, target, source, rng=rng)
        bursts.append(("data", DMR_DT_VOICE_LC, lc))
        emb_lcs = [lc[:9]]
        if with_ta:
            # talker alias header + block 1 (reference src/dmr_decoder/talkeralias.cpp:58-144): the 7 payload
            # bytes start at LC byte 2; header byte = format(2) | length(5) | 1 spare bit
            name = b"B200 TESTER"
            hdr = [4, 0, (1 << 6) | (len(name) << 1)] + list(name[:6])             # 8-bit format
            emb_lcs.append(hdr[:9])
            rest = list(name[6:13])
            blk1 = [5, 0] + rest + [0] * (7 - len(rest))
            emb_lcs.append(blk1[:9])
        if with_gps:
            emb_lcs.append([8, 0, 0x01, 0x12, 0x34, 0x56, 0x23, 0x45, 0x67])
        for s in range(n_super):
            frags = dmr_embedded_lc_fragments(emb_lcs[s % len(emb_lcs)])
            bursts.append(("voice", None, None))
            for q, lcss in enumerate([1, 3, 3, 2]):
                bursts.append(("voice", (1, 0, lcss), frags[q]))
            bursts.append(("voice", (1, 0, 0), None))
        bursts.append(("data", DMR_DT_TERMINATOR_LC, lc))

    if kind == "idle":
        for _ in range(40):
            bursts.append(("data", DMR_DT_IDLE, list(rng.integers(0, 256, size=12))))
    elif kind == "data":
        for i in range(40):
            dt = [DMR_DT_CSBK, DMR_DT_IDLE, DMR_DT_RATE_34, 6, 7][i % 5]
            bursts.append(("data", dt, list(rng.integers(0, 256, size=12))))
    elif kind == "voice":
        voice_call(6, int(rng.integers(1, 1 << 24)), int(rng.integers(1, 1 << 24)), group=True, with_ta=True,
                   with_gps=True)
    else:
        for _ in range(3):
            bursts.append(("data", DMR_DT_IDLE, list(rng.integers(0, 256, size=12))))
        voice_call(2, int(rng.integers(1, 1 << 24)), int(rng.integers(1, 1 << 24)), group=bool(rng.integers(0, 2)))
        for _ in range(4):
            bursts.append(("data", DMR_DT_CSBK, list(rng.integers(0, 256, size=12))))
        voice_call(3, int(rng.integers(1, 1 << 24)), int(rng.integers(1, 1 << 24)), group=False, with_ta=True)
    return bursts


def dmr_symbols(n_frames, seed=0, kinds=("voice", "mixed"), lead_in=None, symbol_errors=0.0):
    """Dibit stream of a DMR base-station carrier: the two slots alternate, each following its own script.
    lead_in random dibits precede the first frame; symbol_errors is the per-dibit probability of a random error."""
    rng = np.random.default_rng(seed)
    scripts = [dmr_slot_script(rng, kinds[0]), dmr_slot_script(rng, kinds[1])]
    idx = [0, 0]
    if lead_in is None:
        lead_in = int(rng.integers(0, 200))
    out = [rng.integers(0, 4, size=lead_in).astype(np.uint8)]
    for fno in range(n_frames):
        slot = fno & 1
        sc = scripts[slot]
        kind, a, b = sc[idx[slot] % len(sc)]
        idx[slot] += 1
        if kind == "data":
            out.append(dmr_data_burst(slot, a, b, rng))
        else:
            out.append(dmr_voice_burst(slot, rng, emb=a, fragment=b))
    s = np.concatenate(out)
    if symbol_errors > 0:
        hit = rng.random(s.size) < symbol_errors
        s = np.where(hit, s ^ rng.integers(1, 4, size=s.size).astype(np.uint8), s).astype(np.uint8)
    return s


# ---------------------------------------------------------------------------------------------------------------
# Batched modulation with torch (harness glue): the same rectangular-hold model as modulate(), vectorised over
# channels and runnable on the GPU so that bench-sized inputs (thousands of channels) are built in seconds.
def modulate_batch(symbols, n_samples, sps=10, levels=LEVELS4, amplitude=0.5, ppm=0.0, phase=0.0, snr_db=None,
                   dc=0.0, seed=0, device="cpu", pitch=None, chunk_channels=512):
    """symbols: uint8 array [C, S].  Per-channel scalars may be arrays of length C (snr_db: inf = no noise).
    Returns a float32 torch tensor [C, pitch] on `device` (pitch >= n_samples, padding zero)."""
    import torch
    symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
    C, S = symbols.shape
    if pitch is None:
        pitch = (n_samples + 3) & ~3

    def vec(v):
        return torch.as_tensor(np.broadcast_to(np.asarray(v, dtype=np.float64), (C,)).copy(), device=device)

    amp, ppm_t, ph, dc_t = vec(amplitude), vec(ppm), vec(phase), vec(dc)
    snr = vec(np.inf if snr_db is None else snr_db)
    lev = torch.as_tensor(np.asarray(levels, dtype=np.float64), device=device)
    p_sig = float(np.mean(np.asarray(levels) ** 2))
    out = torch.zeros((C, pitch), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    t = torch.arange(n_samples, dtype=torch.float64, device=device)
    sym_t = torch.as_tensor(symbols, device=device)
    for c0 in range(0, C, chunk_channels):
        c1 = min(C, c0 + chunk_channels)
        idx = torch.floor((t[None, :] + ph[c0:c1, None]) * (1.0 + ppm_t[c0:c1, None] * 1e-6) / sps).to(torch.int64)
        idx.clamp_(0, S - 1)
        x = lev[torch.gather(sym_t[c0:c1], 1, idx).to(torch.int64)] * amp[c0:c1, None] + dc_t[c0:c1, None]
        sigma = torch.sqrt(p_sig * amp[c0:c1] ** 2 / torch.pow(10.0, snr[c0:c1] / 10.0))
        sigma = torch.where(torch.isfinite(snr[c0:c1]), sigma, torch.zeros_like(sigma))
        noise = torch.randn((c1 - c0, n_samples), dtype=torch.float32, device=device, generator=gen)
        x = x + noise.to(torch.float64) * sigma[:, None]
        out[c0:c1, :n_samples] = x.clamp_(-1.0, 1.0).to(torch.float32)
    return out


def dmr_channel_bank(channels, n_samples, seed=0, device="cpu", pool=48, noise_fraction=0.1):
    """Synthetic workload of SURVEY.md §8d C2: `channels` 48 kHz channels carrying DMR base-station traffic
    (voice superframes, LC headers/terminators, idle/CSBK/rate-3/4 bursts, valid FEC everywhere), with per-channel
    start phase, amplitude, DC offset, AWGN (SNR from {inf, 20, 12, 6} dB), sampling-clock offset
    (0, +-20, +-50 ppm) and ~10 % channels of pure noise.  Returns (float32 tensor [channels, pitch], info dict)."""
    rng = np.random.default_rng(seed)
    sps = 10
    n_sym = n_samples // sps + 300
    frames = n_sym // 144 + 2
    kinds = [("voice", "mixed"), ("mixed", "voice"), ("idle", "voice"), ("data", "mixed"), ("voice", "voice"),
             ("mixed", "data"), ("voice", "idle"), ("mixed", "mixed")]
    base = []
    for k in range(min(pool, channels)):
        s = dmr_symbols(frames, seed=seed * 7919 + k, kinds=kinds[k % len(kinds)], lead_in=0)
        base.append(s[:frames * 144])
    base = np.stack(base)
    S = base.shape[1]
    symbols = np.empty((channels, S), dtype=np.uint8)
    shifts = rng.integers(0, S, size=channels)
    noise_only = rng.random(channels) < noise_fraction
    for c in range(channels):
        if noise_only[c]:
            symbols[c] = rng.integers(0, 4, size=S)
        else:
            symbols[c] = np.roll(base[c % base.shape[0]], int(shifts[c]))
    snr = rng.choice([np.inf, 20.0, 12.0, 6.0], size=channels)
    ppm = rng.choice([0.0, 20.0, -20.0, 50.0, -50.0], size=channels)
    phase = rng.integers(0, 1440, size=channels).astype(np.float64)
    amp = rng.choice([0.25, 0.5, 0.8], size=channels)
    dc = rng.choice([0.0, 0.02, -0.05], size=channels)
    x = modulate_batch(symbols, n_samples, sps=sps, amplitude=amp, ppm=ppm, phase=phase, snr_db=snr, dc=dc,
                       seed=seed + 1, device=device)
    info = {"noise_only": noise_only, "snr_db": snr, "ppm": ppm}
    return x, info


# ---------------------------------------------------------------------------------------------------------------
# POCSAG (reference src/pocsag_decoder/*): preamble, then batches of
# FSC 0x7CD215D8 + 16 codewords; codeword = flag(1) payload(20) BCH(10) even-parity(1), MSB first.
POCSAG_FSC = 0x7CD215D8
POCSAG_IDLE = 0x7A89C197


def pocsag_codeword(flag, payload20):
    data21 = ((flag & 1) << 20) | (payload20 & 0xFFFFF)
    cw31 = bch_31_21_encode(data21)
    cw = cw31 << 1
    return cw | (bin(cw).count("1") & 1)


def pocsag_alpha_payloads(text):
    bits = []
    for ch in text.encode("ascii"):
        bits += [(ch >> j) & 1 for j in range(7)]          # LSB first per character
    while len(bits) % 20:
        bits.append(0)
    return [int("".join(str(b) for b in bits[i:i + 20]), 2) for i in range(0, len(bits), 20)]


def pocsag_bits(messages, seed=0, preamble=576, bit_errors=0, trailing_batches=1, lead_in=0):
    """messages: list of (address, function, text).  Returns the bit stream (uint8 0/1) the demodulator should
    emit.  bit_errors random bit flips are injected into every codeword (<= 2 are correctable)."""
    rng = np.random.default_rng(seed)
    words = []          # list of batches, each 16 codewords

    def new_batch():
        words.append([POCSAG_IDLE] * 16)

    new_batch()
    slot = 0
    for address, function, text in messages:
        frame = address & 7
        start = frame * 2
        if slot > start:
            new_batch()
            slot = 0
        slot = start
        queue = [pocsag_codeword(0, ((address >> 3) << 2) | (function & 3))]
        queue += [pocsag_codeword(1, p) for p in pocsag_alpha_payloads(text)]
        for cw in queue:
            if slot >= 16:
                new_batch()
                slot = 0
            words[-1][slot] = cw
            slot += 1
        if slot >= 16:
            new_batch()
            slot = 0
        else:
            slot += 1       # at least one idle codeword terminates the message
    for _ in range(trailing_batches):
        new_batch()
    out = [rng.integers(0, 2, size=lead_in).astype(np.uint8)] if lead_in else []
    out.append((np.arange(preamble) % 2 == 0).astype(np.uint8))
    for batch in words:
        out.append(_int_to_bits(POCSAG_FSC, 32))
        for cw in batch:
            b = _int_to_bits(cw, 32)
            if bit_errors:
                for e in rng.choice(32, size=bit_errors, replace=False):
                    b[e] ^= 1
            out.append(b)
    return np.concatenate(out).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------
# YSF (reference src/ysf_decoder/*): 480-dibit frames = sync(20) FICH(100) payload(360)
YSF_SYNC = _hex_to_dibits("D471C9634D")


def crc16_ccitt(data):
    """crc16_checksum of the reference (src/ysf_decoder/crc16.c:3-19): poly 0x1021, init 0, final inversion."""
    crc = 0
    for byte in data:
        for i in range(8):
            inp = (int(byte) >> (7 - i)) & 1
            nx = inp ^ ((crc >> 15) & 1)
            crc = (crc << 1) & 0xFFFF
            crc ^= (nx << 12) | (nx << 5) | nx
    return crc ^ 0xFFFF


def pn9_bits(n):
    """Whitening sequence of src/ysf_decoder/whitening.c:7-20."""
    wsr = 0b111001001
    out = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        wb = wsr & 1
        out[i] = wb
        fb = ((wsr >> 4) & 1) ^ wb
        wsr = ((wsr & 0b111111110) >> 1) | (fb << 8)
    return out


def ysf_conv_encode(bits):
    """Rate-1/2 K=5 encoder matching trellis_transitions (src/ysf_decoder/trellis.c:8-25): state = last four
    input bits, newest at the MSB; returns one dibit per input bit."""
    state = 0
    out = np.zeros(len(bits), dtype=np.uint8)
    for i, b in enumerate(bits):
        t = 3 if b else 0
        if state & 1:
            t ^= 3
        if state & 2:
            t ^= 2
        if state & 4:
            t ^= 1
        if state & 8:
            t ^= 1
        out[i] = t
        state = (int(b) << 3) | (state >> 1)
    return out


def ysf_fich_dibits(fi, fn, dt, rng):
    """FICH: FI at bits 31-30, FN at 21-19, DT at 9-8 (src/ysf_decoder/fich.cpp:56-66); other bits random."""
    fich = int(rng.integers(0, 1 << 32))
    fich &= ~((3 << 30) | (7 << 19) | (3 << 8))
    fich |= (fi << 30) | (fn << 19) | (dt << 8)
    be = [(fich >> 24) & 0xFF, (fich >> 16) & 0xFF, (fich >> 8) & 0xFF, fich & 0xFF]
    crc = crc16_ccitt(be)
    bits48 = _int_to_bits((fich << 16) | crc, 48)
    coded = []
    for i in range(4):
        data12 = int("".join(str(b) for b in bits48[12 * i:12 * i + 12]), 2)
        coded.append(_int_to_bits(encode_block("golay_24_12", data12), 24))
    bits = np.concatenate(coded + [np.zeros(4, dtype=np.uint8)])
    enc = ysf_conv_encode(bits)
    tx = np.zeros(100, dtype=np.uint8)
    for i in range(100):
        tx[(i * 20) % 100 + (i * 20) // 100] = enc[i]
    return tx


def _ysf_dch_encode(data, rng):
    """data: 10 or 20 bytes -> conv-encoded dibits (100 or 180): whiten, CRC16 over the whitened bytes, 4 tail bits."""
    data = np.asarray(data, dtype=np.uint8)
    nbits = data.size * 8
    w = np.packbits(np.unpackbits(data) ^ pn9_bits(nbits))
    crc = crc16_ccitt(w)
    bits = np.concatenate([np.unpackbits(w), _int_to_bits(crc, 16), np.zeros(4, dtype=np.uint8)])
    return ysf_conv_encode(bits)


def _callsign(text):
    b = text.encode("latin-1")[:10]
    return np.frombuffer(b + b" " * (10 - len(b)), dtype=np.uint8)


def ysf_gps_frames(lat_digits=(4, 8, 0, 7, 3, 0), south=False, west=False, lon_c=0x30, lon_min=0x40, lon_frac=0x30,
                   radio=0x28, lon_hi=0x30):
    """DT1 + DT2 (2 x 10 bytes) of a short-GPS data frame (src/ysf_decoder/data.cpp:34-88, gps.cpp:7-105)."""
    d = np.zeros(20, dtype=np.uint8)
    d[1:4] = [0x22, 0x62, 0x5F]
    d[4] = radio
    g = [0x30 | v for v in lat_digits]
    g[3] = (0x30 if south else 0x50) | lat_digits[3]
    g[4] = lon_hi | lat_digits[4]
    g[5] = (0x50 if west else 0x30) | lat_digits[5]
    d[5:11] = g
    d[11] = lon_c
    d[12] = lon_min
    d[13] = lon_frac
    d[18] = 0x03
    d[19] = int(d[:19].sum()) & 0xFF
    return d[:10], d[10:]


def ysf_frame(fi, fn, dt, rng, fields=None, gps=None):
    """One 480-dibit frame.  fields: dict with 'dest','src','down','up' callsigns."""
    fields = fields or {}
    f = np.zeros(480, dtype=np.uint8)
    f[0:20] = YSF_SYNC
    f[20:120] = ysf_fich_dibits(fi, fn, dt, rng)
    payload = rng.integers(0, 4, size=360).astype(np.uint8)
    if fi in (0, 2):
        for which, (a, b) in enumerate([("dest", "src"), ("down", "up")]):
            csd = np.concatenate([_callsign(fields.get(a, "")), _callsign(fields.get(b, ""))])
            enc = _ysf_dch_encode(csd, rng)
            for i in range(180):
                streampos = (i % 9) * 20 + i // 9
                payload[(streampos // 36) * 72 + streampos % 36 + 36 * which] = enc[i]
    elif dt == 2:
        if fn <= 3:
            data = _callsign(fields.get(["dest", "src", "down", "up"][fn], ""))
        elif fn >= 6 and gps is not None:
            data = gps[fn - 6]
        else:
            data = rng.integers(0, 256, size=10).astype(np.uint8)
        enc = _ysf_dch_encode(data, rng)
        for i in range(100):
            payload[(i % 5) * 72 + i // 5] = enc[i]
    f[120:480] = payload
    return f


def ysf_symbols(n_frames, seed=0, mode="DN", lead_in=None, symbol_errors=0.0):
    """A YSF transmission: header, n communication frames of the given mode ('DN' = V/D2, 'V1', 'VW' = voice FR,
    'FR' = data FR, 'mix'), terminator, then noise; repeated until n_frames frames exist."""
    rng = np.random.default_rng(seed)
    calls = ["DL1ABC", "ALL", "RPT-DOWN", "RPT-UP", "JA1YSF", "W1AW", "B200 TEST", "X"]
    if lead_in is None:
        lead_in = int(rng.integers(0, 300))
    out = [rng.integers(0, 4, size=lead_in).astype(np.uint8)]
    made = 0
    while made < n_frames:
        fields = {"dest": calls[int(rng.integers(0, 8))], "src": calls[int(rng.integers(0, 8))],
                  "down": calls[int(rng.integers(0, 8))], "up": calls[int(rng.integers(0, 8))]}
        gps = ysf_gps_frames(lat_digits=tuple(int(v) for v in rng.integers(0, 6, size=6)),
                             south=bool(rng.integers(0, 2)), west=bool(rng.integers(0, 2)),
                             lon_c=int(rng.integers(0x26, 0x7f)), lon_min=int(rng.integers(0x26, 0x58)),
                             lon_frac=int(rng.integers(0x1c, 0x7f)))
        m = mode if mode != "mix" else ["DN", "V1", "VW", "FR"][int(rng.integers(0, 4))]
        dt = {"V1": 0, "FR": 1, "DN": 2, "VW": 3}[m]
        out.append(ysf_frame(0, 0, dt, rng, fields))
        made += 1
        n_comm = int(rng.integers(6, 20))
        for k in range(n_comm):
            out.append(ysf_frame(1, k % 8, dt, rng, fields, gps))
            made += 1
        out.append(ysf_frame(2, 0, dt, rng, fields))
        made += 1
        out.append(rng.integers(0, 4, size=int(rng.integers(0, 700))).astype(np.uint8))
    s = np.concatenate(out)
    if symbol_errors > 0:
        hit = rng.random(s.size) < symbol_errors
        s = np.where(hit, s ^ rng.integers(1, 4, size=s.size).astype(np.uint8), s).astype(np.uint8)
    return s


# ---------------------------------------------------------------------------------------------------------------
# NXDN (reference src/nxdn_decoder/*): 192-dibit frames = FSW(10) + LICH(8) + SACCH(30) + 2 x 72 (voice pairs or
# FACCH1); everything behind the FSW is scrambled by inverting the symbol when the PN9 output is 1.
NXDN_FSW = np.array([3, 0, 3, 1, 3, 3, 1, 1, 2, 1], dtype=np.uint8)   # nxdn_phase.cpp:17


def nxdn_scramble(dibits):
    """Scrambler of src/nxdn_decoder/scrambler.cpp:12-25 (its own inverse): register 0b011100100."""
    sr = 0b011100100
    out = np.array(dibits, dtype=np.uint8)
    for i in range(out.size):
        wb = sr & 1
        out[i] = (out[i] & 3) ^ (wb << 1)
        fb = ((sr >> 4) & 1) ^ wb
        sr = ((sr & 0b111111110) >> 1) | (fb << 8)
    return out


def nxdn_crc(bits, width, poly, init):
    """Bit-serial CRC of sacch.cpp:76-90 (width 6, poly 0b010011) / facch1.cpp:60-75 (width 12, poly 0b10000000111)."""
    crc = init
    top = width - 1
    mask = (1 << width) - 1
    for b in bits:
        cb = ((crc >> top) & 1) ^ int(b)
        if cb:
            crc ^= poly
        crc = ((crc << 1) & (mask & ~1)) | cb
    return crc


def _nxdn_channel_encode(info_bits, crc_width, crc_poly, crc_init, punct, rows, cols):
    """info + CRC + 4 tail zeros -> rate-1/2 K=5 code (same generator as YSF) -> puncture -> block interleave.
    punct(i) is True for coded-bit positions that are NOT transmitted; the receiver reads tx[i * cols + k] as
    punctured bit k * rows + i (sacch.cpp:45-54, facch1.cpp:34-43).  Returns dibits."""
    info_bits = np.asarray(info_bits, dtype=np.uint8)
    crc = nxdn_crc(info_bits, crc_width, crc_poly, crc_init)
    bits = np.concatenate([info_bits, _int_to_bits(crc, crc_width), np.zeros(4, dtype=np.uint8)])
    enc = ysf_conv_encode(bits)
    coded = np.empty(2 * enc.size, dtype=np.uint8)
    coded[0::2] = enc >> 1
    coded[1::2] = enc & 1
    kept = np.array([coded[i] for i in range(coded.size) if not punct(i)], dtype=np.uint8)
    assert kept.size == rows * cols
    tx = np.empty_like(kept)
    for i in range(rows):
        for k in range(cols):
            tx[i * cols + k] = kept[k * rows + i]
    return (tx[0::2] << 1) | tx[1::2]


def nxdn_sacch_dibits(structure, ran, data18):
    """structure: 0..3 = position in the superframe (SR field = structure ^ 3, sacch.cpp:18-20)."""
    info = np.concatenate([_int_to_bits(structure ^ 3, 2), _int_to_bits(ran, 6), np.asarray(data18, dtype=np.uint8)])
    return _nxdn_channel_encode(info, 6, 0b010011, 0b111111, lambda i: (i + 1) % 6 == 0, 12, 5)


def nxdn_facch1_dibits(data80):
    return _nxdn_channel_encode(data80, 12, 0b10000000111, 0xFFF, lambda i: (i - 1) % 4 == 0, 16, 9)


def nxdn_lich_dibits(rf, functional, option, direction, rng, bad_parity=False):
    """8 dibits: the LICH bit is the HIGH bit of each dibit (lich.cpp:10-13); low bits are 1 on air (+-3 symbols)."""
    v = (rf << 5) | (functional << 3) | (option << 1) | direction
    bits = _int_to_bits(v, 7)
    par = int(bits[0] ^ bits[1] ^ bits[2] ^ bits[3]) ^ int(bad_parity)
    bits = np.concatenate([bits, [par]]).astype(np.uint8)
    return (bits << 1) | 1


def nxdn_frame(lich, sacch, halves, rng):
    """lich: 8 dibits, sacch: 30 dibits, halves: two arrays of 72 dibits -> one scrambled 192-dibit frame."""
    body = np.concatenate([lich, sacch, halves[0], halves[1]]).astype(np.uint8)
    return np.concatenate([NXDN_FSW, nxdn_scramble(body)])


def nxdn_symbols(n_frames, seed=0, lead_in=None, symbol_errors=0.0):
    """NXDN traffic: calls made of voice frames (SACCH superframes carrying VCALL with call type / source /
    destination), FACCH1-stolen halves (IDLE and other message types), frames with broken LICH parity,
    non-superframe SACCH, UDCH and RCCH frames, ended by a TX_RELEASE FACCH1; noise gaps in between."""
    rng = np.random.default_rng(seed)
    if lead_in is None:
        lead_in = int(rng.integers(0, 400))
    out = [rng.integers(0, 4, size=lead_in).astype(np.uint8)]
    made = 0

    def facch(msg_type):
        d = rng.integers(0, 2, size=80).astype(np.uint8)
        d[2:8] = _int_to_bits(msg_type, 6)
        return nxdn_facch1_dibits(d)

    while made < n_frames:
        call_type = int(rng.choice([0b001, 0b100, 0b000, 0b110]))
        src, dst = int(rng.integers(1, 65536)), int(rng.integers(0, 65536))
        ran = int(rng.integers(0, 64))
        msg = int(rng.choice([0x01, 0x01, 0x01, 0x07]))            # mostly VCALL
        sf = np.concatenate([_int_to_bits(int(rng.integers(0, 4)), 2), _int_to_bits(msg, 6),
                             _int_to_bits(int(rng.integers(0, 256)), 8), _int_to_bits(call_type, 3),
                             _int_to_bits(int(rng.integers(0, 32)), 5), _int_to_bits(src, 16), _int_to_bits(dst, 16),
                             _int_to_bits(int(rng.integers(0, 65536)), 16)])
        n_call = int(rng.integers(8, 40))
        for k in range(n_call):
            structure = k % 4
            r = rng.random()
            rf, functional, option = 0b10, 0b10, 0b11
            if r < 0.08:
                option = int(rng.integers(0, 3))                   # one or both halves stolen by FACCH1
            elif r < 0.11:
                functional = 0b00                                  # non-superframe SACCH: not collected
            elif r < 0.13:
                functional = 0b01                                  # UDCH: frame skipped
            elif r < 0.15:
                rf = 0b00                                          # RCCH: frame skipped
            lich = nxdn_lich_dibits(rf, functional, option, int(rng.integers(0, 2)), rng,
                                    bad_parity=rng.random() < 0.05)
            sacch = nxdn_sacch_dibits(structure, ran, sf[18 * structure:18 * structure + 18])
            halves = []
            for i in range(2):
                if (option >> (1 - i)) & 1:
                    halves.append(rng.integers(0, 4, size=72).astype(np.uint8))
                else:
                    halves.append(facch(int(rng.choice([0x10, 0x10, 0x01, 0x3F]))))
            out.append(nxdn_frame(lich, sacch, halves, rng))
            made += 1
        # end of call: TX_RELEASE in the first or second half
        which = int(rng.integers(0, 2))
        halves = [facch(0x08) if which == 0 else rng.integers(0, 4, size=72).astype(np.uint8),
                  facch(0x08) if which == 1 else facch(0x10)]
        option = 0b00 if which == 0 else 0b10
        lich = nxdn_lich_dibits(0b10, 0b10, option, 0, rng)
        out.append(nxdn_frame(lich, nxdn_sacch_dibits(n_call % 4, ran, sf[0:18]), halves, rng))
        made += 1
        out.append(rng.integers(0, 4, size=int(rng.integers(0, 500))).astype(np.uint8))
    s = np.concatenate(out)
    if symbol_errors > 0:
        hit = rng.random(s.size) < symbol_errors
        s = np.where(hit, s ^ rng.integers(1, 4, size=s.size).astype(np.uint8), s).astype(np.uint8)
    return s


# ---------------------------------------------------------------------------------------------------------------
# D-Star (reference src/dstar_decoder/*): 4800 bit/s, one bit per symbol.  Transmission = bit sync, frame sync,
# 660-bit radio header (K=3 rate-1/2 code, 24-row interleave, scrambler), then 96-bit frames of 72 voice bits + 24
# slow-data bits (every 21st data field is the voice sync), closed by the 48-bit terminator.
DSTAR_HEADER_SYNC = np.array([0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0, 0], dtype=np.uint8)
DSTAR_VOICE_SYNC = np.array([1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0], dtype=np.uint8)
DSTAR_TERMINATOR = np.array([1, 0] * 16 + [0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 0], dtype=np.uint8)


def dstar_scramble(bits):
    """src/dstar_decoder/scrambler.cpp:6-21 (its own inverse): register 0b1111111, output bit0 ^ bit3."""
    sr = 0b1111111
    out = np.array(bits, dtype=np.uint8) & 1
    for i in range(out.size):
        wb = (sr & 1) ^ ((sr >> 3) & 1)
        out[i] ^= wb
        sr = ((sr & 0b1111110) >> 1) | (wb << 6)
    return out


def dstar_crc(data):
    """src/dstar_decoder/crc.cpp:6-23: reflected CCITT (0x8408), init 0xFFFF, inverted."""
    c = 0xFFFF
    for byte in data:
        for i in range(8):
            c ^= (int(byte) >> i) & 1
            c = (c >> 1) ^ 0x8408 if c & 1 else c >> 1
    return c ^ 0xFFFF


def _lsb_bits(data):
    return np.unpackbits(np.asarray(data, dtype=np.uint8), bitorder="little")


def dstar_header_bytes(flags=(0, 0, 0), rpt2="DB0XYZ G", rpt1="DB0XYZ B", your="CQCQCQ", my="DL1ABC", suffix="B200"):
    def f(t, n):
        b = t.encode("latin-1")[:n]
        return b + b" " * (n - len(b))
    body = bytes(flags) + f(rpt2, 8) + f(rpt1, 8) + f(your, 8) + f(my, 8) + f(suffix, 4)
    crc = dstar_crc(body)
    return np.frombuffer(body + bytes([crc & 0xFF, crc >> 8]), dtype=np.uint8)


def dstar_header_bits(header41):
    """41 bytes -> 660 channel bits (header.cpp:23-48 backwards: bits LSB first, K=3 code of
    Header::trellis_transitions, interleave, scramble)."""
    bits = np.concatenate([_lsb_bits(header41), np.zeros(2, dtype=np.uint8)])
    state = 0
    coded = np.zeros(660, dtype=np.uint8)
    for i, b in enumerate(bits):
        t = 3 if b else 0
        if state & 1:
            t ^= 3
        if state & 2:
            t ^= 2
        coded[2 * i] = t >> 1
        coded[2 * i + 1] = t & 1
        state = (int(b) << 1) | (state >> 1)
    tx = np.zeros(660, dtype=np.uint8)
    for i in range(12):
        for k in range(28):
            tx[i * 28 + k] = coded[k * 24 + i]
    for i in range(12, 24):
        for k in range(27):
            tx[12 + i * 27 + k] = coded[k * 24 + i]
    return dstar_scramble(tx)


def dstar_slow_data_blocks(message=None, header41=None, text=None):
    """6-byte slow-data blocks (mini header + 5 bytes, dstar_phase.cpp:163-211)."""
    blocks = []
    if message is not None:
        m = message.encode("latin-1")[:20]
        m = m + b" " * (20 - len(m))
        for k in range(4):
            blocks.append(bytes([0x40 | k]) + m[5 * k:5 * k + 5])
    if header41 is not None:
        h = bytes(header41)
        for k in range(0, 41, 5):
            part = h[k:k + 5]
            blocks.append(bytes([0x50 | len(part)]) + part + b"\x66" * (5 - len(part)))
    if text is not None:
        t = text.encode("latin-1")
        for k in range(0, len(t), 5):
            part = t[k:k + 5]
            blocks.append(bytes([0x30 | len(part)]) + part + b"\x66" * (5 - len(part)))
    return blocks


def dstar_gga(lat=4807.038, lon=1131.0, south=False, west=False):
    body = "GPGGA,123519,%09.4f,%s,%010.4f,%s,1,08,0.9,545.4,M,46.9,M,," % (lat, "S" if south else "N", lon,
                                                                              "W" if west else "E")
    cs = 0
    for ch in body:
        cs ^= ord(ch)
    return "$%s*%02X\r\n" % (body, cs)


def dstar_dprs(payload="DL1ABC-7>API282,DSTAR*:!4807.03N/01131.00E>B200 test"):
    body = payload + "\r"
    return "$$CRC%04X,%s" % (dstar_crc(body.encode("latin-1")), body)


def dstar_symbols(n_frames, seed=0, lead_in=None, bit_errors=0.0, gga=True):
    """D-Star traffic: transmissions with radio header (voice, some data headers, some with a broken CRC), voice
    frames carrying slow data (20-character message, header resend, DPRS and NMEA GGA sentences, filler), late
    entry (voice sync without header), terminators (full and second half only), noise gaps."""
    rng = np.random.default_rng(seed)
    calls = ["DL1ABC", "W1AW", "JA1YSF", "OE1XYZ", "G4KLX"]
    if lead_in is None:
        lead_in = int(rng.integers(0, 300))
    out = [rng.integers(0, 2, size=lead_in).astype(np.uint8)]
    made = 0
    while made < n_frames:
        my = calls[int(rng.integers(0, len(calls)))]
        flags = (0x80 if rng.random() < 0.1 else 0x00, 0, 0)
        hdr = dstar_header_bytes(flags=flags, my=my, your=calls[int(rng.integers(0, len(calls)))],
                                 suffix=["B200", "", "ID51"][int(rng.integers(0, 3))])
        late_entry = rng.random() < 0.2
        if not late_entry:
            hb = dstar_header_bits(hdr)
            if rng.random() < 0.1:
                hb = hb.copy()
                hb[rng.choice(660, size=40, replace=False)] ^= 1      # uncorrectable header
            out.append(np.concatenate([np.tile([1, 0], 32).astype(np.uint8), DSTAR_HEADER_SYNC[9:], hb]))
        blocks = []
        r = rng.random()
        if r < 0.5:
            blocks += dstar_slow_data_blocks(message="B200 msg %d %s" % (made, my))
        if r > 0.3:
            blocks += dstar_slow_data_blocks(header41=hdr)
        # gga=False keeps the random draws but leaves the sentence out: the REFERENCE aborts (std::stof throws) on a
        # GGA sentence whose fields were corrupted by bit errors while its 8-bit checksum still matches
        if rng.random() < 0.6:
            gga_blocks = dstar_slow_data_blocks(text=dstar_gga(lat=float(rng.uniform(0, 8959)), lon=float(rng.uniform(0, 17959)),
                                                               south=bool(rng.integers(0, 2)), west=bool(rng.integers(0, 2))))
            if gga:
                blocks += gga_blocks
        if rng.random() < 0.6:
            blocks += dstar_slow_data_blocks(text=dstar_dprs("%s>API282,DSTAR*:!4807.03N/01131.00E>run %d" % (my, made)))
        filler = bytes([0x66] * 6)
        n_voice = int(rng.integers(25, 90))
        halves = []
        for b in blocks:
            halves += [b[:3], b[3:]]
        hi = 0
        for k in range(n_voice):
            voice = rng.integers(0, 2, size=72).astype(np.uint8)
            sync_slot = (k % 21 == 0)
            if sync_slot:
                data = DSTAR_VOICE_SYNC
            else:
                h3 = halves[hi] if hi < len(halves) else filler[:3]
                hi += 1
                data = dstar_scramble(_lsb_bits(np.frombuffer(h3, dtype=np.uint8)))
            out.append(np.concatenate([voice, data]))
            made += 1
        voice = rng.integers(0, 2, size=72).astype(np.uint8)
        if rng.random() < 0.7:
            out.append(np.concatenate([voice, DSTAR_TERMINATOR]))
        else:
            out.append(np.concatenate([voice, DSTAR_TERMINATOR[24:]]))
        made += 1
        out.append(rng.integers(0, 2, size=int(rng.integers(0, 600))).astype(np.uint8))
    s = np.concatenate(out)
    if bit_errors > 0:
        s = (s ^ (rng.random(s.size) < bit_errors)).astype(np.uint8)
    return s
