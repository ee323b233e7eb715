"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the CPU legs of bench.py.

Everything here runs the UNMODIFIED reference (oracle/_ref/libria_ref.so, oracle/_ref/cli_simulator; the C
port oracle/libria_oracle.so only where the reference library is missing) on the host cores: bench.py's
`cpu_baseline` object and its `--impl reference` arm.  Nothing in this module imports ria_b200 -- inputs are
synthesised with the reference's own transmitter -- so a process that only runs the reference arm never loads
the product library.

Each `cpu_*` function takes the workload parameters as a plain dict (bench.py owns the configuration) and
returns {"value", "unit", "cores", "kind", "sample"}.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import re
import subprocess
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CLI = os.path.join(HERE, "_ref", "cli_simulator")


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _fan(worker, jobs):
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        return pool.map(worker, jobs, chunksize=1)


def _awgn(tx, snr_db, rng):
    p = float(np.mean(tx.astype(np.float64) ** 2))
    return (tx + rng.standard_normal(len(tx)).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr_db / 10)))).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# OFDM fixed frames: processPresynced + complete decodeFixedFrame + parseHeader
# ---------------------------------------------------------------------------------------------
_POOLS = {}


def ofdm_pool(p):
    """p["pool"] clean TX frames made by the reference transmitter (makeData -> encodeFixedFrame ->
    OFDMModulator), no chunk 1..3 starting with 0xD5 (CodewordStatus::reassemble would drop such frames)."""
    span = int(p.get("cfo_span", 0))
    cox = p.get("cox")                                       # {"lead": .., "window": ..}: Schmidl-Cox preamble, window rows
    key = ("ofdm", p["modulation"], p["pilot_spacing"], p["rate"], p["pool"], span, str(cox))
    if key not in _POOLS:
        from oracle.bindings import BITS_PER_CARRIER, BYTES_PER_CW, ModemConfig, Ref
        ref = Ref()
        cfg = ModemConfig.make(p["modulation"], p["pilot_spacing"], 1)
        bps = cfg.data_carriers() * BITS_PER_CARRIER[cfg.modulation]
        bpc = BYTES_PER_CW[p["rate"]]
        rng = np.random.default_rng(11)
        pool, offs = [], []
        while len(pool) < p["pool"]:
            frame = ref.make_data_frame("K1ABC", "W2XYZ", len(pool), rng.integers(0, 256, size=4 * bpc - 19, dtype=np.uint8))
            if any(frame[c * bpc] == 0xD5 for c in (1, 2, 3)):
                continue
            # a transmitter whose carrier is off by d Hz is the frequency-shifted signal (CFO variant)
            d = int(rng.integers(-span, span + 1)) if span else 0
            tx_cfg = ModemConfig.from_buffer_copy(bytes(cfg))
            tx_cfg.center_freq = cfg.center_freq + d
            coded = ref.encode_fixed_frame(frame, p["rate"], True, bps)
            if cox:
                tx = ref.ofdm_cox_tx_frame(tx_cfg, coded)
                row = np.zeros(cox["window"], np.float32)
                row[cox["lead"]:cox["lead"] + len(tx)] = tx[: cox["window"] - cox["lead"]]
                pool.append(row)
            else:
                pool.append(ref.ofdm_tx_frame(tx_cfg, coded))
            offs.append(d)
        _POOLS[key] = (cfg, bps, pool, offs)
    return _POOLS[key]


def _ofdm_worker(args):
    p, n_frames, seed = args
    from oracle.bindings import Ref
    ref = Ref()
    cfg, bps, pool, offs = ofdm_pool(p)
    rng = np.random.default_rng(seed)
    frames = [_awgn(pool[i % len(pool)], p["snr_db"], rng) for i in range(n_frames)]
    # the receiver is told the frame's carrier offset the way a sync stage would: +- 0.2 Hz
    cfo = np.array([offs[i % len(pool)] + (rng.uniform(-0.2, 0.2) if p.get("cfo_span") else 0.0) for i in range(n_frames)])
    ok = 0
    cox = p.get("cox")
    t0 = time.perf_counter()
    for rx, c in zip(frames, cfo):
        if cox:                                              # OFDMNvisWaveform::detectSync + process
            found, pos, c, _ = ref.ofdm_cox_search_sync(cfg, rx, 0.8, 0.0)
            if not found:
                continue
            ph = np.float32(-2.0 * np.pi * float(c) * float(pos) / 48000.0)
            while float(ph) > np.pi:
                ph = np.float32(float(ph) - 2.0 * np.pi)
            while float(ph) < -np.pi:
                ph = np.float32(float(ph) + 2.0 * np.pi)
            r = ref.ofdm_process_presynced(cfg, rx[pos:pos + cox["frame_len"]], float(c), float(ph))
            data, cw_ok = ref.decode_fixed_frame_full(r["soft"], p["rate"], True, bps)
            ok += int(cw_ok.all() and ref.parse_header(data).frame_crc_ok)
            continue
        r = ref.ofdm_process_presynced(cfg, rx, float(c), 0.0)
        data, cw_ok = ref.decode_fixed_frame_full(r["soft"], p["rate"], True, bps)
        ok += int(cw_ok.all() and ref.parse_header(data).frame_crc_ok)
    return time.perf_counter() - t0, ok


def cpu_ofdm(p, frames_per_core):
    from oracle.bindings import Ref
    if not Ref.available():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libria_ref.so not present"}
    cores = host_cores()
    ofdm_pool(p)                               # built once, inherited by the forked workers
    t0 = time.perf_counter()
    res = _fan(_ofdm_worker, [(p, frames_per_core, 1000 + 17 * c) for c in range(cores)])
    wall = time.perf_counter() - t0
    rate = sum(frames_per_core / b for b, _ in res)
    ok = sum(o for _, o in res)
    return {"value": rate, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{frames_per_core} frames per core x {cores} cores ({ok}/{frames_per_core * cores} decoded with valid CRC), "
                      f"reference {'searchForSync (Schmidl-Cox acquisition) + ' if p.get('cox') else ''}processPresynced + complete decodeFixedFrame + parseHeader, one process per core, "
                      f"{wall:.1f} s wall, {max(b for b, _ in res):.1f} s max busy"}


# ---------------------------------------------------------------------------------------------
# MC-DPSK: detectDualChirp + MultiCarrierDPSKDemodulator::process + LDPCDecoder::decodeSoft
# ---------------------------------------------------------------------------------------------
def mcdpsk_pool(p):
    key = ("mcdpsk", p["pool"], p["lead"], p["tail"])
    if key not in _POOLS:
        from oracle.bindings import McdpskConfig, Ref
        ref = Ref()
        cfg = McdpskConfig.make(1, 4, 10)
        rng = np.random.default_rng(31)
        pre = ref.chirp_generate()
        rows = []
        frame_len = 0
        for _ in range(p["pool"]):
            coded = ref.ldpc_encode(p["rate"], rng.integers(0, 256, size=20, dtype=np.uint8))[:81]
            body = ref.mcdpsk_tx_frame(cfg, coded)
            frame_len = len(body)
            rows.append(np.concatenate([np.zeros(p["lead"], np.float32), pre, body, np.zeros(p["tail"], np.float32)]))
        _POOLS[key] = (cfg, frame_len, rows)
    return _POOLS[key]


def _mcdpsk_worker(args):
    p, n_frames, seed = args
    from oracle.bindings import Ref
    ref = Ref()
    cfg, frame_len, pool = mcdpsk_pool(p)
    rng = np.random.default_rng(seed)
    rows = [_awgn(pool[i % len(pool)], p["snr_db"], rng) for i in range(n_frames)]
    ok = 0
    t0 = time.perf_counter()
    for rx in rows:
        s = ref.chirp_detect_dual(rx[:p["window"]], 0.15)
        if not s.detected:
            continue
        start = int(s.aux) + 28800
        r = ref.mcdpsk_process(cfg, rx[start:start + frame_len], float(s.cfo_hz))
        if len(r["soft"]) >= 648:
            _, okk, _ = ref.ldpc_decode_batch(p["rate"], r["soft"][:648], p["max_iter"], p["factor"], 24)
            ok += int(okk[0])
    return time.perf_counter() - t0, ok


def cpu_mcdpsk(p, frames_per_core):
    from oracle.bindings import Ref
    if not Ref.available():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libria_ref.so not present"}
    cores = host_cores()
    mcdpsk_pool(p)
    t0 = time.perf_counter()
    res = _fan(_mcdpsk_worker, [(p, frames_per_core, 2000 + 13 * c) for c in range(cores)])
    wall = time.perf_counter() - t0
    rate = sum(frames_per_core / b for b, _ in res)
    ok = sum(o for _, o in res)
    return {"value": rate, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{frames_per_core} single receptions per core x {cores} cores ({ok}/{frames_per_core * cores} decoded), "
                      f"reference detectDualChirp + MC-DPSK process + LDPC decodeSoft, one process per core, {wall:.1f} s wall; "
                      f"the GPU figure counts TWO receptions per frame"}


# ---------------------------------------------------------------------------------------------
# LDPC decodeSoft on synthetic AWGN soft bits
# ---------------------------------------------------------------------------------------------
def ldpc_llr_host(p, rate, n, seed, impl):
    rng = np.random.default_rng(seed)
    nbytes = {0: 20, 1: 27, 2: 40, 3: 54, 4: 60}.get(rate, 20)
    cws = np.stack([np.unpackbits(np.asarray(impl.ldpc_encode(rate, rng.integers(0, 256, size=nbytes, dtype=np.uint8)))[:81])
                    for _ in range(32)])
    snr = np.float32(10 ** (p["esn0"][rate] / 10))
    s = 1.0 - 2.0 * cws[rng.integers(0, len(cws), size=n)].astype(np.float32)
    noise = rng.standard_normal(s.shape, dtype=np.float32) / np.sqrt(snr)
    return (2.0 * (s + noise) * snr).astype(np.float32)


def _ldpc_worker(args):
    p, n, seed, kind = args
    from oracle.bindings import Port, Ref
    impl = Ref() if kind == "reference" else Port()
    elapsed = 0.0
    for rate in p["rates"]:
        llr = ldpc_llr_host(p, rate, n, seed + rate, impl)
        t0 = time.perf_counter()
        impl.ldpc_decode_batch(rate, llr, p["max_iter"][rate], p["factor"])
        elapsed += time.perf_counter() - t0
    return elapsed


def cpu_ldpc(p, per_rate_per_core):
    from oracle.bindings import Ref
    kind = "reference" if Ref.available() else "port"
    cores = host_cores()
    t0 = time.perf_counter()
    busy = _fan(_ldpc_worker, [(p, per_rate_per_core, 1000 + 17 * c, kind) for c in range(cores)])
    wall = time.perf_counter() - t0
    per_core_cw = len(p["rates"]) * per_rate_per_core
    rate_sum = sum(per_core_cw / b for b in busy)
    return {"value": rate_sum / p["cw_per_frame"], "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{per_rate_per_core} codewords per rate per core x {cores} cores ({cores * per_core_cw} codewords), "
                      f"{wall:.1f} s wall, {max(busy):.1f} s max busy"}


# ---------------------------------------------------------------------------------------------
# cli_simulator, one instance per host core (what BASELINE.json's north_star names)
# ---------------------------------------------------------------------------------------------
def cli_simulator_per_core(overfeed: int = 100, timeout_s: int = 240):
    """One reference cli_simulator process per host core, OFDM DQPSK R1/2 on AWGN 15 dB (BASELINE configs[0]),
    all started together.  cli_simulator runs the complete two-station protocol (connect, data, ACKs,
    disconnect) with four threads per process and paces its audio loop against air time
    (tools/cli_simulator.cpp:1333-1337); --rx-overfeed-factor shortens the wall time.  Rate = frames the two
    stations decoded, summed over the instances, per second of wall time."""
    if not os.path.exists(CLI):
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/cli_simulator not built"}
    cores = host_cores()
    args = [CLI, "--snr", "15", "--channel", "awgn", "-w", "ofdm_chirp", "-m", "dqpsk", "-r", "r1_2",
            "--rx-overfeed-factor", str(overfeed)]
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        procs = [subprocess.Popen(args + ["--seed", str(100 + c)], cwd=tmp, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True) for c in range(cores)]
        decoded = passed = 0
        for pr in procs:
            try:
                out, _ = pr.communicate(timeout=timeout_s)
            except subprocess.TimeoutExpired:
                pr.kill()
                out, _ = pr.communicate()
            decoded += sum(int(m) for m in re.findall(r"frames_decoded=(\d+)", out))
            passed += int("TEST PASSED" in out)
        wall = time.perf_counter() - t0
    return {"value": decoded / wall, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{cores} cli_simulator instances (one per core, 4 threads each; --snr 15 --channel awgn -w ofdm_chirp "
                      f"-m dqpsk -r r1_2 --rx-overfeed-factor {overfeed}), {passed}/{cores} sessions passed, {decoded} frames "
                      f"decoded by the two stations of all instances in {wall:.1f} s wall (complete protocol session, paced "
                      f"against air time: not a compute-bound figure)"}
