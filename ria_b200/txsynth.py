"""Host-side TX synthesis (numpy) used to produce synthetic receive-chain inputs.

Follows the reference transmitter closely enough that its output decodes under the reference
receiver (checked in tests/test_txsynth_cpu.py):
  DataFrame::makeData / serialize      src/protocol/frame_v2.cpp:459-545, hashCallsign :78-84
  v2::encodeFixedFrame                 src/protocol/frame_v2.cpp:1285-1328
  LDPCEncoder (systematic, H=[Hd|I])   src/fec/ldpc_encoder.cpp:193-257
  ChannelInterleaver / FrameInterleaver src/fec/ldpc_decoder.cpp:579-615, frame_interleaver.cpp:37-94
  OFDMModulator::generateTrainingSymbols / modulate / mapBits
                                       src/ofdm/modulator.cpp:20-120, 199-283, 348-477, 528-582
The IFFT runs in float64 numpy, so samples equal the reference TX to ~1e-7, not bit for bit; an
on-device sample-identical TX is a "next" row.
"""
from __future__ import annotations

import numpy as np

from . import fec
from .ofdm import (BPSK, DBPSK, DQPSK, QAM16, QAM32, QAM64, QPSK, ModemConfig, channel_interleaver_step,
                   crc16, getBitsPerSymbol)


def hash_callsign(cs: str) -> int:
    h = 5381
    for ch in cs.upper():
        h = (((h << 5) + h) ^ ord(ch)) & 0xFFFFFFFF
    return h & 0xFFFFFF


def make_data_frame(src: str, dst: str, seq: int, payload: bytes, ftype: int = 0x30) -> bytes:
    payload = bytes(payload)
    total_bits = (17 + len(payload) + 2) * 8
    total_cw = ((total_bits + 161) // 162) & 0xFF           # calculateCodewords (R1/4 basis)
    sh, dh = hash_callsign(src), hash_callsign(dst)
    hdr = bytes([0x55, 0x4C, ftype, 0x01, (seq >> 8) & 0xFF, seq & 0xFF,
                 (sh >> 16) & 0xFF, (sh >> 8) & 0xFF, sh & 0xFF,
                 (dh >> 16) & 0xFF, (dh >> 8) & 0xFF, dh & 0xFF,
                 total_cw, (len(payload) >> 8) & 0xFF, len(payload) & 0xFF])
    h = crc16(hdr)
    body = hdr + bytes([h >> 8, h & 0xFF]) + payload
    f = crc16(body)
    return body + bytes([f >> 8, f & 0xFF])


def _crc16_rows(rows: np.ndarray) -> np.ndarray:
    """CRC-16/CCITT-FALSE of every row of a uint8 [n, len] array (table driven, vectorised over rows)."""
    tab = np.zeros(256, np.uint16)
    for b in range(256):
        r = b << 8
        for _ in range(8):
            r = ((r << 1) ^ 0x1021) & 0xFFFF if r & 0x8000 else (r << 1) & 0xFFFF
        tab[b] = r
    crc = np.full(rows.shape[0], 0xFFFF, np.uint16)
    for j in range(rows.shape[1]):
        crc = ((crc << 8) & 0xFFFF).astype(np.uint16) ^ tab[((crc >> 8) ^ rows[:, j]) & 0xFF]
    return crc


def make_data_frames(src: str, dst: str, seq0: int, payloads: np.ndarray, bytes_per_cw: int = 0) -> np.ndarray:
    """make_data_frame for every row of payloads (uint8 [n, len]) -> frames uint8 [n, 17 + len + 2], seq = seq0 + row.
    With bytes_per_cw > 0 a payload byte that would make a codeword chunk 1..3 start with 0xD5 is changed to
    0xD4 first (see make_frame_pool: the reference drops such frames)."""
    payloads = np.ascontiguousarray(payloads, np.uint8)
    n, ln = payloads.shape
    out = np.zeros((n, 17 + ln + 2), np.uint8)
    sh, dh = hash_callsign(src), hash_callsign(dst)
    seq = (seq0 + np.arange(n)) & 0xFFFF
    total_cw = (((17 + ln + 2) * 8 + 161) // 162) & 0xFF
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = 0x55, 0x4C, 0x30, 0x01
    out[:, 4], out[:, 5] = seq >> 8, seq & 0xFF
    out[:, 6:9] = [(sh >> 16) & 0xFF, (sh >> 8) & 0xFF, sh & 0xFF]
    out[:, 9:12] = [(dh >> 16) & 0xFF, (dh >> 8) & 0xFF, dh & 0xFF]
    out[:, 12], out[:, 13], out[:, 14] = total_cw, (ln >> 8) & 0xFF, ln & 0xFF
    h = _crc16_rows(out[:, :15])
    out[:, 15], out[:, 16] = h >> 8, h & 0xFF
    out[:, 17:17 + ln] = payloads
    if bytes_per_cw:
        for c in (1, 2, 3):
            pos = c * bytes_per_cw
            if 17 <= pos < 17 + ln:
                col = out[:, pos]
                col[col == 0xD5] = 0xD4
    f = _crc16_rows(out[:, :17 + ln])
    out[:, 17 + ln], out[:, 18 + ln] = f >> 8, f & 0xFF
    return out


_H_CACHE = {}


def _h_rows(rate: int):
    if rate not in _H_CACHE:
        k, m, _ = fec.code_params(rate)
        row_ptr, edge_var = fec.get_matrix(rate)
        _H_CACHE[rate] = (k, m, [edge_var[row_ptr[i]:row_ptr[i + 1] - 1] for i in range(m)])
    return _H_CACHE[rate]


def ldpc_encode_bits(info_bits: np.ndarray, rate: int) -> np.ndarray:
    """info_bits [..., k] -> codeword bits [..., 648] (systematic)."""
    k, m, rows = _h_rows(rate)
    info_bits = np.asarray(info_bits, np.uint8)
    par = np.stack([info_bits[..., r].sum(axis=-1) & 1 for r in rows], axis=-1).astype(np.uint8)
    return np.concatenate([info_bits, par], axis=-1)


def encode_fixed_frame_bits(frame: bytes, rate: int, use_channel_interleave: bool, bits_per_symbol: int):
    """-> 2592 interleaved coded bits of a 4-codeword frame."""
    k, _, _ = fec.code_params(rate)
    bpc = k // 8
    data = np.zeros(4 * bpc, np.uint8)
    fb = np.frombuffer(frame[: 4 * bpc], np.uint8)
    data[: len(fb)] = fb
    cws = np.zeros((4, 648), np.uint8)
    for c in range(4):
        info = np.zeros(k, np.uint8)
        info[: bpc * 8] = np.unpackbits(data[c * bpc:(c + 1) * bpc])
        cw = ldpc_encode_bits(info, rate)
        if use_channel_interleave:
            step = channel_interleaver_step(bits_per_symbol, 648)
            out = np.zeros(648, np.uint8)
            out[(np.arange(648) * step) % 648] = cw          # interleaved[perm[i]] = bits[i]
            cw = out
        cws[c] = cw
    inter = np.zeros(2592, np.uint8)
    b = np.arange(648)
    for c in range(4):
        inter[b * 4 + (c + b) % 4] = cws[c]
    return inter


def _map_bits(bits: np.ndarray, mod: int) -> np.ndarray:
    """mapBits (modulator.cpp:73-118) for coherent constellations."""
    if mod == BPSK:
        return np.where(bits & 1, 1.0, -1.0) + 0j
    if mod == QPSK:
        s = 0.7071067811865476
        return np.where(bits & 2, s, -s) + 1j * np.where(bits & 1, s, -s)
    if mod == QAM16:
        lv = np.array([-3, -1, 3, 1]) * 0.3162277660168379
        return lv[(bits >> 2) & 3] + 1j * lv[bits & 3]
    if mod == QAM32:
        s = 0.1961161351381840
        il = np.zeros(4); ql = np.zeros(8)
        for i, g in enumerate([0, 1, 3, 2]):
            il[g] = [-3, -1, 1, 3][i] * s
        for i, g in enumerate([0, 1, 3, 2, 6, 7, 5, 4]):
            ql[g] = [-7, -5, -3, -1, 1, 3, 5, 7][i] * s
        return il[bits & 3] + 1j * ql[(bits >> 2) & 7]
    if mod == QAM64:
        lv = np.array([-7, -5, -1, -3, 7, 5, 1, 3]) * 0.1543033499620919
        return lv[(bits >> 3) & 7] + 1j * lv[bits & 7]
    raise ValueError("unsupported modulation for txsynth")


def _carriers(cfg: ModemConfig):
    nc, N = cfg.num_carriers, cfg.fft_size
    data, pilots = [], []
    idx = 0
    for i in range(-(nc // 2), (nc + 1) // 2 + 1):
        if i == 0:
            continue
        f = (i + N) % N
        if cfg.use_pilots and idx % cfg.pilot_spacing == 0:
            pilots.append(f)
        else:
            data.append(f)
        idx += 1
    return np.array(data), np.array(pilots, dtype=int)


def _mt19937_bits(n: int, seed: int) -> np.ndarray:
    """first n outputs of std::mt19937(seed), lowest bit (pilot signs, demodulator.cpp:91-94)."""
    mt = [0] * 624
    mt[0] = seed
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
    for i in range(624):
        y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
    out = []
    for i in range(n):
        y = mt[i]
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        out.append(y & 1)
    return np.array(out, dtype=int)


_NCO_CACHE = {}


def nco_phasors(center_freq: float, sample_rate: float, n: int) -> np.ndarray:
    """NCO::next sequence (src/dsp/filters.cpp:228-238): fp32 phase accumulator, double-promoted
    wrap.  TX and RX run the same recurrence, so its rounding drift cancels end to end."""
    key = (center_freq, sample_rate)
    have = _NCO_CACHE.get(key)
    if have is None or len(have) < n:
        inc = np.float32(2.0 * np.pi * np.float64(np.float32(center_freq)) / np.float64(np.float32(sample_rate)))
        two_pi = 2.0 * np.pi
        ph = np.float32(0.0)
        phases = np.empty(max(n, 32768), np.float32)
        for i in range(len(phases)):
            phases[i] = ph
            ph = np.float32(ph + inc)
            if float(ph) > two_pi:
                ph = np.float32(float(ph) - two_pi)
            if float(ph) < 0:
                ph = np.float32(float(ph) + two_pi)
        have = (np.cos(phases.astype(np.float64)) + 1j * np.sin(phases.astype(np.float64)))
        _NCO_CACHE[key] = have
    return have[:n]


def ofdm_modulate_frame(cfg: ModemConfig, coded_bits: np.ndarray, output_scale: float = 40.0) -> np.ndarray:
    """2 LTS + data symbols, real passband fp32 samples (phase-continuous mixer from 0)."""
    N = cfg.fft_size
    L = cfg.getSymbolDuration()
    cp = L - N - cfg.symbol_guard
    data_idx, pilot_idx = _carriers(cfg)
    nd = len(data_idx)
    mod = cfg.modulation
    bpc = getBitsPerSymbol(mod)
    n = np.arange(cfg.num_carriers)
    zc = np.exp(1j * (-np.pi * n * (n + 1) / cfg.num_carriers))
    pil = np.where(_mt19937_bits(len(pilot_idx), 0x50494C54) == 1, 1.0, -1.0) + 0j
    bits = np.asarray(coded_bits, np.uint8)
    n_sym = (len(bits) + nd * bpc - 1) // (nd * bpc)
    syms = []
    lts = zc[:nd]
    for _ in range(2):
        syms.append(lts)
    prev = np.ones(nd, complex)
    pos = 0
    for _ in range(n_sym):
        vals = np.zeros(nd, complex)
        for c in range(nd):
            if pos >= len(bits):
                break                                     # remaining carriers stay zero (:455-458)
            chunk = bits[pos:pos + bpc]
            pos += len(chunk)
            v = 0
            for b in chunk:
                v = (v << 1) | int(b)
            v <<= bpc - len(chunk)
            if mod == DQPSK:
                prev[c] = prev[c] * [1, 1j, -1, -1j][v & 3]
                vals[c] = prev[c]
            elif mod == DBPSK:
                prev[c] = prev[c] * (-1 if v & 1 else 1)
                vals[c] = prev[c]
            else:
                vals[c] = _map_bits(np.array(v), mod)
        syms.append(vals)
    out = np.zeros(len(syms) * L, np.float64)
    osc = nco_phasors(cfg.center_freq, cfg.sample_rate, len(out))
    for s, vals in enumerate(syms):
        fd = np.zeros(N, complex)
        fd[data_idx] = vals
        if len(pilot_idx):
            fd[pilot_idx] = pil
        td = np.fft.ifft(fd)                               # 1/N scaling like FFT::inverse
        sym = np.concatenate([td[N - cp:], td])
        seg = slice(s * L, s * L + cp + N)
        out[seg] = np.real(sym * osc[seg]) * output_scale
    return out.astype(np.float32)


def make_frame_pool(cfg: ModemConfig, rate: int, n_frames: int, seed: int = 1,
                    use_channel_interleave: bool = True):
    """n_frames distinct clean TX frames -> (samples fp32 [n, frame_len], frame bytes list).

    Payloads are re-drawn until none of the codeword chunks 1..3 of the frame starts with 0xD5: the
    reference's CodewordStatus::reassemble (src/protocol/frame_v2.cpp:974) takes such a chunk for a
    DATA_CW_MARKER, and its complete decodeFixedFrame then drops the frame however clean the channel
    (tests/test_ldpc_retry_gpu.py pins that behaviour); a throughput workload should not contain them."""
    rng = np.random.default_rng(seed)
    k, _, _ = fec.code_params(rate)
    bpc = k // 8
    bps = cfg.getDataCarriers() * getBitsPerSymbol(cfg.modulation)
    frames, raw = [], []
    for i in range(n_frames):
        while True:
            payload = rng.integers(0, 256, size=4 * bpc - 19, dtype=np.uint8).tobytes()
            fr = make_data_frame("K1ABC", "W2XYZ", i & 0xFFFF, payload)
            if all(fr[c * bpc] != 0xD5 for c in (1, 2, 3) if c * bpc < len(fr)):
                break
        bits = encode_fixed_frame_bits(fr, rate, use_channel_interleave, bps)
        frames.append(ofdm_modulate_frame(cfg, bits))
        raw.append(fr)
    return np.stack(frames), raw


# ---------------------------------------------------------------------------------------------
# Preambles and the MC-DPSK modulator (inputs of the C3 bench workload and of the sync tests)
# ---------------------------------------------------------------------------------------------
F32 = np.float32


def zc_preamble(root: int = 5, sample_rate: float = 48000.0, sequence_length: int = 127, upsample: int = 8,
                repetitions: int = 2, carrier_freq: float = 1500.0, gap_ms: float = 10.0) -> np.ndarray:
    """sync::ZCSync::generatePreambleForRoot (src/sync/zc_sync.hpp:133-190): linearly interpolated
    Zadoff-Chu chips on a 1500 Hz carrier, peak-normalised to 0.8, followed by a silent gap.
    Root 5 = DATA frames (mc_dpsk_waveform.cpp:57-61)."""
    n = np.arange(sequence_length, dtype=np.float64)
    if sequence_length % 2 == 0:
        phase = (-np.pi * root * n * n / sequence_length).astype(F32)
    else:
        phase = (-np.pi * root * n * (n + 1) / sequence_length).astype(F32)
    zc = (np.cos(phase).astype(F32) + 1j * np.sin(phase).astype(F32)).astype(np.complex64)
    rep_len = sequence_length * upsample
    i = np.arange(rep_len)
    chip_pos = (i.astype(F32) / F32(upsample)).astype(F32)
    chip_idx = chip_pos.astype(np.int64)
    frac = (chip_pos - chip_idx.astype(F32)).astype(F32)
    nxt = np.minimum(chip_idx + 1, sequence_length - 1)
    interp = np.where(chip_idx < sequence_length - 1,
                      zc[chip_idx] * (F32(1.0) - frac) + zc[nxt] * frac, zc[chip_idx]).astype(np.complex64)
    gi = (np.arange(repetitions)[:, None] * rep_len + i[None, :]).reshape(-1)
    t = (gi.astype(F32) / F32(sample_rate)).astype(F32)
    cph = (F32(2.0) * F32(np.pi) * F32(carrier_freq) * t).astype(F32)
    it = np.tile(interp, repetitions)
    s = (it.real * np.cos(cph).astype(F32) - it.imag * np.sin(cph).astype(F32)).astype(F32)
    peak = np.abs(s).max()
    if peak > 0:
        s = (s * F32(F32(0.8) / peak)).astype(F32)
    gap = int(F32(sample_rate) * F32(gap_ms) / F32(1000.0))
    return np.concatenate([s, np.zeros(gap, F32)])


def chirp_preamble(sample_rate: float = 48000.0, f_start: float = 300.0, f_end: float = 2700.0,
                   duration_ms: float = 500.0, gap_ms: float = 100.0, amplitude: float = 0.5) -> np.ndarray:
    """sync::ChirpSync::generate, dual chirp (src/sync/chirp_sync.hpp:61-108): [up][gap][down][gap]."""
    n = int(F32(sample_rate) * F32(duration_ms) / F32(1000.0))
    gap = int(F32(sample_rate) * F32(gap_ms) / F32(1000.0))
    T = F32(duration_ms) / F32(1000.0)
    k = F32((F32(f_end) - F32(f_start)) / T)
    t = (np.arange(n).astype(F32) / F32(sample_rate)).astype(F32)
    up = (F32(2.0) * F32(np.pi) * (F32(f_start) * t + F32(0.5) * k * t * t)).astype(F32)
    dn = (F32(2.0) * F32(np.pi) * (F32(f_end) * t - F32(0.5) * k * t * t)).astype(F32)
    out = np.zeros(2 * n + 2 * gap, F32)
    out[:n] = F32(amplitude) * np.sin(up).astype(F32)
    out[n + gap: 2 * n + gap] = F32(amplitude) * np.sin(dn).astype(F32)
    return out


def _mcdpsk_carrier_table(cfg) -> np.ndarray:
    """e^{j 2 pi f_c i / fs}, i < samples_per_symbol, per carrier (multi_carrier_dpsk.hpp:68-79, 155-160)."""
    C, sps = int(cfg.num_carriers), int(cfg.samples_per_symbol)
    if C == 1:
        freqs = np.array([(F32(cfg.freq_low) + F32(cfg.freq_high)) / F32(2.0)], F32)
    else:
        spacing = F32((F32(cfg.freq_high) - F32(cfg.freq_low)) / F32(C - 1))
        freqs = (F32(cfg.freq_low) + np.arange(C).astype(F32) * spacing).astype(F32)
    inc = (F32(2.0) * F32(np.pi) * freqs / F32(cfg.sample_rate)).astype(F32)
    t = (np.arange(sps).astype(F32)[None, :] * inc[:, None]).astype(F32)
    return (np.cos(t).astype(F32) + 1j * np.sin(t).astype(F32)).astype(np.complex64)


def mcdpsk_modulate_frame(cfg, data: bytes) -> np.ndarray:
    """[training][reference][data] as MultiCarrierDPSKModulator emits them
    (src/psk/multi_carrier_dpsk.hpp:141-275): every carrier restarts at phase 0 each symbol, the
    symbol is the mean over carriers of Re(sym_c * e^{j w_c i}), data symbols are repeated
    `spreading` times, DBPSK 0 -> 0, 1 -> pi; DQPSK 00/01/11/10 -> +45/+135/-135/-45 degrees."""
    C, sps, bps = int(cfg.num_carriers), int(cfg.samples_per_symbol), int(cfg.bits_per_symbol)
    spread, n_train = int(cfg.spreading), int(cfg.training_symbols)
    car = _mcdpsk_carrier_table(cfg)                               # [C][sps]

    def symbol(vals: np.ndarray) -> np.ndarray:
        out = np.zeros(sps, F32)
        for c in range(C):                                          # accumulation order of the reference
            out = (out + ((vals[c] * car[c]).real.astype(F32) / F32(C)).astype(F32)).astype(F32)
        return out

    parts = []
    for sym in range(n_train):
        ph = (np.arange(C) * sym).astype(F32) * F32(np.pi) / F32(2.0)
        parts.append(symbol((np.cos(ph).astype(F32) + 1j * np.sin(ph).astype(F32)).astype(np.complex64)))
    prev = np.ones(C, np.complex64)
    parts.append(symbol(prev))
    bits = np.unpackbits(np.frombuffer(bytes(data), np.uint8))
    per_sym = C * bps
    n_ds = (len(bits) + per_sym - 1) // per_sym
    bits = np.concatenate([bits, np.zeros(n_ds * per_sym - len(bits), np.uint8)])
    dq = np.array([np.pi / 4, 3 * np.pi / 4, -3 * np.pi / 4, -np.pi / 4], F32)
    for d in range(n_ds):
        b = bits[d * per_sym:(d + 1) * per_sym].reshape(C, bps)
        if bps == 2:
            change = dq[b[:, 0] * 2 + b[:, 1]]
        else:
            change = np.where(b[:, 0] == 1, F32(np.pi), F32(0.0)).astype(F32)
        diff = (np.cos(change).astype(F32) + 1j * np.sin(change).astype(F32)).astype(np.complex64)
        cur = (prev * diff).astype(np.complex64)
        cur = (cur / np.abs(cur).astype(F32)).astype(np.complex64)
        prev = cur
        s = symbol(cur)
        parts.extend([s] * spread)
    return np.concatenate(parts)
