#!/usr/bin/env python
"""bench.py -- the measurement contract (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one pass of the receive hot path over one batch of synthetic received frames that
are already resident in HBM.  Rank 0 prints ONE JSON line.  For N > 1 the driver launches this
file under torch.distributed.run (one rank per GPU, NCCL); frames are sharded across ranks (weak
scaling: fixed per-GPU batch, no data-path collective) and the only exchange is one all-reduce of
the error counters.

Workloads
  ofdm_qam64   (default) BASELINE.json configs[3]: OFDM QAM64 R3/4 coherent, 15 pilots, 1024-FFT
               CP 96, AWGN 28 dB, 1M frames per GPU: presynced demod -> de-interleave -> 4 x LDPC
               -> header/CRC, i.e. "decoded frames/s".
  ldpc         BASELINE.json configs[1]: LDPC R1/4..R3/4 decode of 1M codewords per rate.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KK_LDPC, KK_OFDM_DEMOD, KK_FRAME_STATUS, KK_AWGN = 0, 1, 2, 3
KK_OFDM_FFT, KK_OFDM_CARRIER, KK_OFDM_PHASE = 11, 12, 13
KK_LDPC_RETRY, KK_FRAME_REPAIR = 14, 15
KK_MCDPSK, KK_CHIRP_SYNC, KK_CHASE, KK_MCDPSK_CFO = 4, 6, 7, 9


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------
# workload: OFDM QAM64 R3/4 full receive chain (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------

class OfdmQam64Workload:
    name = "ofdm_qam64_r34_15pilots_awgn28dB"
    metric = "decoded_frames_per_s"
    unit = "frames/s"
    dtype = "f32"
    MOD, RATE, SPACING, SNR_DB = 8, 4, 4, 28.0          # QAM64, R3/4, pilot every 4th carrier
    POOL = 64

    def __init__(self, n_frames: int):
        self.n = n_frames
        self._pool = None
        self.frame_len = 12 * 1120

    def cfg(self):
        from ria_b200 import ofdm
        return ofdm.ModemConfig.high_throughput(self.MOD)

    def pool_host(self, seed=11):
        """POOL distinct clean TX frames (host numpy TX synthesiser) and their frame bytes."""
        if self._pool is None:
            from ria_b200 import txsynth
            self._pool = txsynth.make_frame_pool(self.cfg(), self.RATE, self.POOL, seed=seed)
            self.frame_len = self._pool[0].shape[1]
        return self._pool

    def describe(self):
        c = self.cfg()
        L = c.getSymbolDuration()
        return {"workload": self.name, "frames_per_gpu": self.n, "frame_samples": self.frame_len,
                "modulation": "QAM64", "code_rate": "R3/4", "fft": 1024, "cp": L - 1024,
                "carriers": 59, "pilots": c.getPilotCarriers(), "data_carriers": c.getDataCarriers(),
                "channel": f"AWGN {self.SNR_DB} dB generated on the device (Philox); every frame has its own payload, "
                           "encodeFixedFrame + OFDM TX on the device (sample-identical to the reference transmitter)",
                "chain": "mix+CFO+CP+FFT1024+LTS/pilot est+MMSE+QAM64 LLR -> frame/channel "
                         "de-interleave -> complete v2::decodeFixedFrame (4x LDPC R3/4 0.9375/60 it, retry ladder and "
                         "false-positive repair armed: RIA_DECODE_FULL) -> header+CRC16",
                "l2": "input batch (53.8 GB at 1M frames) exceeds the 126 MB L2; no flush needed"}

    def setup(self, ctx, device, rank, world):
        import torch
        from ria_b200 import ofdm, sim
        self.torch, self.ctx = torch, ctx
        # Every frame is its own transmission: payload bytes (host RNG) -> v2 data frame (header, CRCs) ->
        # encodeFixedFrame, OFDM training + modulate and the AWGN channel on the device
        # (ria_encode_fixed_frame_batch_dev / ria_ofdm_tx_frames_dev are byte- / sample-identical to the
        # reference transmitter, tests/test_tx_gpu.py).  Generated in chunks so that only one chunk of clean
        # TX samples exists next to the resident RX batch.
        from ria_b200 import txsynth
        bpc = 60
        self.first_id = rank * self.n                       # global frame ids: sharding-independent payloads and noise
        cfg = self.cfg()
        bps = cfg.getDataCarriers() * 6
        self.samples = torch.empty((self.n, self.frame_len), dtype=torch.float32, device=device)
        self.sent_len = 4 * bpc - 2
        self.sent_dev = torch.empty((self.n, self.sent_len), dtype=torch.uint8, device=device)
        chunk = 65536
        for off in range(0, self.n, chunk):
            m = min(chunk, self.n - off)
            rng = np.random.default_rng([2026, self.first_id + off])
            frames = txsynth.make_data_frames("K1ABC", "W2XYZ", self.first_id + off,
                                              rng.integers(0, 256, size=(m, 4 * bpc - 19 - 2), dtype=np.uint8), bpc)
            fr_dev = torch.from_numpy(frames).to(device)
            self.sent_dev[off:off + m] = fr_dev
            coded = ofdm.encode_fixed_frame_batch(fr_dev, self.RATE, True, bps, ctx)
            tx = ofdm.ofdm_tx_frames(cfg, coded, ctx)
            assert tx.shape[1] == self.frame_len
            # the channel takes TX row (global id % pool rows) and keys its noise by the global id: with the
            # chunk as the pool, rotate it so that frame i of the chunk sits in row (gid + i) % m
            gid = self.first_id + off
            if gid % m:
                tx = torch.roll(tx, shifts=gid % m, dims=0)
            sim.awgn_batch(tx, m, self.SNR_DB, seed=2026, first_frame_id=gid, out=self.samples[off:off + m], ctx=ctx)
            del tx, coded, fr_dev
        import ria_b200
        ctx.set_decode_flags(ria_b200.DECODE_FULL)          # the reference's complete decodeFixedFrame
        self.chain = ofdm.OfdmRxChain(self.cfg(), self.RATE, True, ctx)
        self.out = None
        torch.cuda.synchronize()

    def step(self):
        self.out = self.chain.process_batch(self.samples)

    def launches_per_step(self):
        return 3

    def units_per_step(self):
        return float(self.n)

    def samples_per_step(self):
        return float(self.n) * self.frame_len

    def kernels(self):
        """kind -> (kernel name, ALGORITHMIC bytes per step); DESIGN.md section 3 derives the figures.
        Per frame: 12 symbols x 1120 samples in, 12 x 59 complex bins between the stages, 10 x 44 x 6
        soft bits out of the demodulator, 4 x 61 info bytes + flags out of the decoder."""
        n_sym, nc, n_llr = 12, 59, 10 * 44 * 6
        samples, bins, llr = self.frame_len * 4, n_sym * nc * 8, n_llr * 4
        return {
            KK_OFDM_FFT: ("ofdm_fft_kernel", self.n * (samples + bins)),
            KK_OFDM_CARRIER: ("ofdm_carrier_kernel", self.n * (bins + llr)),
            KK_OFDM_DEMOD: ("ofdm_presynced_kernel (residual-CFO re-run frames only)", 0),
            KK_LDPC: ("ldpc_decode_kernel", self.n * (llr + 4 * (61 + 5))),
            KK_FRAME_STATUS: ("frame_status_kernel", self.n * (4 * 72 + 4 * 5 + 40)),
            KK_LDPC_RETRY: ("ldpc_fail_list_kernel + ldpc_retry_kernel (retry ladder, failed frames only)", self.n * 4),
            KK_FRAME_REPAIR: ("frame_repair_list_kernel + frame_repair_kernel (false-positive repair, invalid frames only)",
                              self.n * (4 + 4 * 60)),
        }

    def counters(self):
        """[frames, frames_ok (4/4 cw + header + frame CRC), cw_fail, frames_payload_wrong, sum_iters]"""
        torch = self.torch
        from ria_b200 import ofdm
        data, status, snr = self.out
        st = status.view(torch.uint8)
        dt = ofdm.FRAME_STATUS_DTYPE
        off = {k: dt.fields[k][1] for k in dt.names}
        cw_ok = st[:, off["cw_ok"]:off["cw_ok"] + 4]
        iters = st[:, off["cw_iters"]:off["cw_iters"] + 16].contiguous().view(torch.int32)
        ok = (st[:, off["all_ok"]] == 1) & (st[:, off["header_valid"]] == 1) & (st[:, off["frame_crc_ok"]] == 1)
        wrong = (data[:, : self.sent_len] != self.sent_dev).any(dim=1) & ok
        c = torch.zeros(5, dtype=torch.int64, device=data.device)
        c[0] = self.n
        c[1] = ok.sum()
        c[2] = (cw_ok == 0).sum()
        c[3] = wrong.sum()
        c[4] = iters.sum()
        return c

    def counter_dict(self, c):
        return {"frames": int(c[0]), "frames_ok": int(c[1]), "cw_fail": int(c[2]),
                "crc_ok_but_payload_wrong": int(c[3]), "mean_ldpc_iters": float(c[4]) / max(1, 4 * int(c[0]))}

    # ---- end to end: HOST buffers through ria_ofdm_rx_frames_host ----
    def setup_e2e(self, n_e2e):
        torch = self.torch
        self.e2e_n = n_e2e = min(n_e2e, self.n)
        pin = torch.empty((n_e2e, self.frame_len), dtype=torch.float32, pin_memory=True)
        pin.copy_(self.samples[:n_e2e])
        torch.cuda.synchronize()
        self._pin = pin
        self.e2e_host = pin.numpy()

    def step_e2e(self):
        return self.chain.process_batch_host(self.e2e_host)

    def e2e_units(self):
        return float(self.e2e_n)

    def e2e_bytes(self):
        from ria_b200 import ofdm
        # ria_ofdm_rx_frames_host copies only the 1024-sample FFT window of each 1120-sample symbol (the
        # cyclic prefix is never read by the demodulator): 12 x 1024 x 4 B per frame cross PCIe
        sym_copied = 1024 if os.environ.get("RIA_H2D_FULL_SYMBOLS") != "1" else 1120
        return (self.e2e_n * (self.frame_len // 1120) * sym_copied * 4,
                self.e2e_n * (240 + ofdm.FRAME_STATUS_DTYPE.itemsize + 4))


def _cpu_ofdm_worker(args):
    """One process per core: make its own received frames (untimed), then time the reference's
    processPresynced + the complete v2::decodeFixedFrame (first pass, retry ladder, false-positive repair) +
    parseHeader on them."""
    _, n_frames, seed, kind = args
    from oracle.bindings import ModemConfig, Ref
    wl = OfdmQam64Workload(n_frames)
    ref = Ref()
    pool, raw = wl.pool_host()
    cfg = ModemConfig.from_buffer_copy(bytes(wl.cfg()))
    rng = np.random.default_rng(seed)
    bps = cfg.data_carriers() * 6
    frames = []
    for i in range(n_frames):
        tx = pool[i % len(pool)]
        p = float(np.mean(tx.astype(np.float64) ** 2))
        frames.append((tx + rng.standard_normal(len(tx)).astype(np.float32) *
                       np.float32(np.sqrt(p / 10 ** (wl.SNR_DB / 10)))).astype(np.float32))
    ok = 0
    t0 = time.perf_counter()
    for rx in frames:
        r = ref.ofdm_process_presynced(cfg, rx, 0.0, 0.0)
        data, cw_ok = ref.decode_fixed_frame_full(r["soft"], wl.RATE, True, bps)
        st = ref.parse_header(data)
        ok += int(cw_ok.all() and st.frame_crc_ok)
    return time.perf_counter() - t0, ok


_OFDM_POOL_CACHE = {}


def cpu_baseline_ofdm(wl, frames_per_core):
    import multiprocessing as mp
    from oracle.bindings import Ref
    if not Ref.available():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/libria_ref.so not present"}
    cores = host_cores()
    jobs = [("ofdm_qam64", frames_per_core, 1000 + 17 * c, "reference") for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_ofdm_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    rate = sum(frames_per_core / b for b, _ in res)
    ok = sum(o for _, o in res)
    return {"value": rate, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{frames_per_core} frames per core x {cores} cores "
                      f"({ok}/{frames_per_core * cores} decoded with valid CRC), reference "
                      f"processPresynced + complete decodeFixedFrame + parseHeader, one process "
                      f"per core, {wall:.1f} s wall, {max(b for b, _ in res):.1f} s max busy"}


# ---------------------------------------------------------------------------------------------
# workload: chirp-acquired MC-DPSK 4x spread with HARQ chase combining (BASELINE.json configs[2])
# ---------------------------------------------------------------------------------------------

class McdpskC3Workload:
    """Every frame is received twice (HARQ): per reception the channel is simulated on the device
    (AWGN at -8 dB over the whole transmission), the dual chirp is searched in the first 120 000
    samples of the row (streaming_decoder.cpp:408-411), the frame is demodulated at the detected
    training start with the detected CFO, its soft bits are chase-combined with the previous
    reception and the LDPC R1/4 codeword is decoded.  A step is both receptions of all frames."""
    name = "mcdpsk_dbpsk_10car_4x_awgn-8dB_chirp_chase"
    metric = "decoded_frames_per_s"
    unit = "frames/s"
    dtype = "f32"
    RATE, MAX_ITER, FACTOR, SNR_DB, POOL, LEAD, TAIL, WINDOW = 0, 50, 0.9375, -8.0, 64, 2000, 800, 120000

    def __init__(self, n_frames: int):
        self.n = n_frames
        self._pool = None
        self.row_len = self.frame_len = 0

    def cfg(self):
        from ria_b200 import mcdpsk
        return mcdpsk.MultiCarrierDPSKConfig.level4_dbpsk(mcdpsk.SPREAD_4X)

    def pool_host(self, seed=31):
        if self._pool is None:
            from ria_b200 import txsynth
            rng = np.random.default_rng(seed)
            pre = txsynth.chirp_preamble()
            rows, sent = [], []
            for _ in range(self.POOL):
                data = rng.integers(0, 256, size=20, dtype=np.uint8)
                info_bits = np.concatenate([np.unpackbits(data), np.zeros(2, np.uint8)])       # k = 162
                cw = np.packbits(txsynth.ldpc_encode_bits(info_bits, self.RATE))
                body = txsynth.mcdpsk_modulate_frame(self.cfg(), cw.tobytes())
                self.frame_len = len(body)
                rows.append(np.concatenate([np.zeros(self.LEAD, np.float32), pre, body, np.zeros(self.TAIL, np.float32)]))
                sent.append(data)
            self._pool = (np.stack(rows), np.stack(sent))
            self.row_len = self._pool[0].shape[1]
        return self._pool

    def describe(self):
        self.pool_host()
        return {"workload": self.name, "frames_per_gpu": self.n, "receptions_per_frame": 2,
                "row_samples": self.row_len, "frame_samples": self.frame_len, "modulation": "DBPSK",
                "carriers": 10, "spreading": 4, "code_rate": "R1/4",
                "channel": f"AWGN {self.SNR_DB} dB over the whole transmission, generated on the device "
                           f"(Philox) inside the timed step, {self.POOL} distinct TX frames",
                "chain": "per reception: AWGN -> dual-chirp sync (131072-pt FFT matched filter) -> Hilbert CFO "
                         "correction -> 10-carrier correlation demod + 4x despreading -> chase combine -> "
                         "LDPC R1/4 (0.9375, 50 it)",
                "l2": "one reception of the batch (78 GB at 100k frames) exceeds the 126 MB L2; no flush needed"}

    def setup(self, ctx, device, rank, world):
        import torch
        from ria_b200 import mcdpsk
        self.torch, self.ctx, self.device = torch, ctx, device
        pool, sent = self.pool_host()
        self.pool_dev = torch.from_numpy(pool).to(device)
        self.sent_dev = torch.from_numpy(sent).to(device)
        self.first_id = rank * self.n
        self.rows = torch.empty((self.n, self.row_len), dtype=torch.float32, device=device)
        self.acc = torch.empty((self.n, 648), dtype=torch.float32, device=device)
        self.chain = mcdpsk.McdpskRxChain(self.cfg(), self.RATE, self.MAX_ITER, self.FACTOR, 0.15, ctx)
        self.out = None
        self.first_ok = None
        self.epoch = 0
        torch.cuda.synchronize()

    def step(self):
        from ria_b200 import sim
        for rec in (0, 1):
            sim.awgn_batch(self.pool_dev, self.n, self.SNR_DB, seed=9000 + 2 * self.epoch + rec,
                           first_frame_id=self.first_id, out=self.rows, ctx=self.ctx)
            self.out = self.chain.process_batch(self.rows, self.frame_len, self.WINDOW, self.acc, rec == 0, self.out)
            if rec == 0:
                self.first_ok = self.out["ok"].clone()
        self.epoch += 1

    def units_per_step(self):
        return float(self.n)

    def samples_per_step(self):
        return 2.0 * self.n * self.row_len

    def extra(self, frames_per_s):
        """every frame is received twice (HARQ round): the chain runs 2 x value receptions per second; the CPU
        baseline counts single receptions"""
        return {"receptions_per_s": 2.0 * frames_per_s}

    def kernels(self):
        """kind -> (kernel names, ALGORITHMIC bytes per step); two receptions per frame."""
        n2 = 2 * self.n
        row, body, win = self.row_len * 4, self.frame_len * 4, self.WINDOW * 4
        return {
            KK_AWGN: ("awgn_kernel", n2 * 2 * row),
            KK_CHIRP_SYNC: ("chirp_* (3 x 3-stage 131072-pt FFT with fused real load / template products, peak)", n2 * win),
            KK_MCDPSK_CFO: ("mcdpsk_phase_scan_kernel + mcdpsk_cfo_kernel", n2 * 2 * body),
            KK_MCDPSK: ("mcdpsk_demod_kernel", n2 * (body + 652 * 4)),
            KK_CHASE: ("chase_combine_kernel", n2 * 3 * 648 * 4),
            KK_LDPC: ("ldpc_decode_kernel", n2 * (648 * 4 + 24 + 5)),
        }

    def counters(self):
        """[frames, ok after the first reception, ok after chase combining, decoded-but-wrong payload, sum iters]"""
        torch = self.torch
        ok2 = self.out["ok"].bool()
        want = self.sent_dev[(torch.arange(self.n, device=self.device) + self.first_id) % self.POOL]
        wrong = (self.out["info"][:, :20] != want).any(dim=1) & ok2
        return torch.stack([torch.tensor(self.n, device=self.device), self.first_ok.sum(), ok2.sum(), wrong.sum(),
                            self.out["iters"].sum()]).to(torch.int64)

    def counter_dict(self, c):
        return {"frames": int(c[0]), "frames_ok_first_reception": int(c[1]), "frames_ok_after_chase_combining": int(c[2]),
                "decoded_but_wrong_payload": int(c[3]), "mean_ldpc_iters_second_pass": float(c[4]) / max(1, int(c[0]))}

    def setup_e2e(self, n_e2e):
        torch = self.torch
        self.e2e_n = n_e2e = min(n_e2e, self.n, 2048)
        pin = torch.empty((n_e2e, self.row_len), dtype=torch.float32, pin_memory=True)
        pin.copy_(self.rows[:n_e2e])
        torch.cuda.synchronize()
        self._pin = pin

    def step_e2e(self):
        return self.chain.process_batch_host(self._pin, self.frame_len, self.WINDOW)

    def e2e_units(self):
        return float(self.e2e_n)

    def e2e_bytes(self):
        return (self.e2e_n * self.row_len * 4, self.e2e_n * (24 + 1 + 4 + 32))


def _cpu_mcdpsk_worker(args):
    """One process per core: reference detectDualChirp + MC-DPSK process + LDPC on its own rows."""
    _, n_frames, seed, _ = args
    from oracle.bindings import McdpskConfig, Ref
    wl = McdpskC3Workload(n_frames)
    ref = Ref()
    pool, _ = wl.pool_host()
    cfg = McdpskConfig.from_buffer_copy(bytes(wl.cfg()))
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n_frames):
        tx = pool[i % len(pool)]
        p = float(np.mean(tx.astype(np.float64) ** 2))
        rows.append((tx + rng.standard_normal(len(tx)).astype(np.float32) *
                     np.float32(np.sqrt(p / 10 ** (wl.SNR_DB / 10)))).astype(np.float32))
    ok = 0
    t0 = time.perf_counter()
    for rx in rows:
        s = ref.chirp_detect_dual(rx[:wl.WINDOW], 0.15)
        if not s.detected:
            continue
        start = int(s.aux) + 28800
        r = ref.mcdpsk_process(cfg, rx[start:start + wl.frame_len], float(s.cfo_hz))
        if len(r["soft"]) >= 648:
            _, okk, _ = ref.ldpc_decode_batch(wl.RATE, r["soft"][:648], wl.MAX_ITER, wl.FACTOR, 24)
            ok += int(okk[0])
    return time.perf_counter() - t0, ok


def cpu_baseline_mcdpsk(wl, frames_per_core):
    import multiprocessing as mp
    from oracle.bindings import Ref
    if not Ref.available():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/libria_ref.so not present"}
    cores = host_cores()
    wl.pool_host()
    jobs = [("mcdpsk_c3", frames_per_core, 2000 + 13 * c, "reference") for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_mcdpsk_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    rate = sum(frames_per_core / b for b, _ in res)
    ok = sum(o for _, o in res)
    return {"value": rate, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{frames_per_core} single receptions per core x {cores} cores ({ok}/{frames_per_core * cores} "
                      f"decoded), reference detectDualChirp + MC-DPSK process + LDPC decodeSoft, one process per "
                      f"core, {wall:.1f} s wall; the GPU figure counts TWO receptions per frame"}


# ---------------------------------------------------------------------------------------------
# workload: batched LDPC decode (BASELINE.json configs[1])
# ---------------------------------------------------------------------------------------------

class LdpcWorkload:
    """configs[1]: R1/4..R3/4 min-sum decode of 1M codewords per rate from synthetic AWGN LLRs
    (LLR model of tools/test_chase_cache.cpp:20-34; Es/N0 and decoder settings of SURVEY.md 8d)."""

    name = "ldpc_r14_r34_1M_cw_per_rate"
    metric = "decoded_frames_per_s"
    unit = "frames/s"
    dtype = "f32"
    RATES = (0, 2, 3, 4)
    ESN0 = {0: 1.0, 2: 4.0, 3: 6.0, 4: 7.0}
    MAX_ITER = {0: 50, 2: 80, 3: 70, 4: 60}
    FACTOR = 0.9375
    CW_PER_FRAME = 4

    def __init__(self, n_cw: int):
        self.n_cw = n_cw

    def describe(self):
        return {"workload": self.name, "codewords_per_rate_per_gpu": self.n_cw,
                "rates": "R1/4,R1/2,R2/3,R3/4", "esn0_db": "1,4,6,7", "min_sum_factor": self.FACTOR,
                "max_iter": "50,80,70,60", "frame": "4 codewords (v2 fixed frame)",
                "l2": "inputs (2.7 GB per rate) exceed the 126 MB L2; no flush needed"}

    def _codeword_bits(self, rate, n_distinct, rng):
        from ria_b200 import fec, txsynth
        k = fec.code_params(rate)[0]
        return txsynth.ldpc_encode_bits(rng.integers(0, 2, size=(n_distinct, k), dtype=np.uint8), rate)

    def make_llr_host(self, rate, n, seed):
        rng = np.random.default_rng(seed)
        bits = self._codeword_bits(rate, 64, rng)
        snr = np.float32(10 ** (self.ESN0[rate] / 10))
        s = 1.0 - 2.0 * bits[rng.integers(0, 64, size=n)].astype(np.float32)
        noise = rng.standard_normal(s.shape, dtype=np.float32) / np.sqrt(snr)
        return (2.0 * (s + noise) * snr).astype(np.float32), bits

    def setup(self, ctx, device, rank, world):
        import torch
        from ria_b200 import fec
        self.torch, self.ctx = torch, ctx
        self.dec, self.llr, self.info_ref, self.prot_mask = {}, {}, {}, {}
        gen = torch.Generator(device=device).manual_seed(1234 + rank)
        for rate in self.RATES:
            rng = np.random.default_rng(99 + rate)
            bits = self._codeword_bits(rate, 256, rng)
            base = torch.from_numpy(bits).to(device)
            pick = torch.randint(0, 256, (self.n_cw,), device=device, generator=gen)
            s = 1.0 - 2.0 * base[pick].float()
            snr = 10 ** (self.ESN0[rate] / 10)
            llr = 2.0 * (s + torch.randn(s.shape, device=device, generator=gen) / snr ** 0.5) * snr
            self.llr[rate] = llr.contiguous()
            k = fec.code_params(rate)[0]
            packed = np.packbits(bits[:, :k], axis=1)
            self.info_ref[rate] = torch.from_numpy(packed).to(device)[pick]
            # info bits that appear in at least one check (R3/4 leaves 162 bits unprotected,
            # SURVEY.md section 7 "Quirks"); errors are counted on protected bits only
            _, edge_var = fec.get_matrix(rate)
            prot = np.zeros(8 * packed.shape[1], np.uint8)
            prot[edge_var[edge_var < k]] = 1
            self.prot_mask[rate] = torch.from_numpy(np.packbits(prot)).to(device)
            d = fec.LDPCDecoder(rate, ctx)
            d.setMaxIterations(self.MAX_ITER[rate])
            d.setMinSumFactor(self.FACTOR)
            self.dec[rate] = d
            del s, pick
        self.out = None
        torch.cuda.synchronize()

    def step(self):
        self.out = {rate: self.dec[rate].decode_batch(self.llr[rate]) for rate in self.RATES}

    def launches_per_step(self):
        return len(self.RATES)

    def units_per_step(self):
        return len(self.RATES) * self.n_cw / self.CW_PER_FRAME

    def samples_per_step(self):
        return 0.0

    def kernels(self):
        from ria_b200 import fec
        return {KK_LDPC: ("ldpc_decode_kernel",
                          sum(self.n_cw * (648 * 4 + (fec.code_params(r)[0] + 7) // 8 + 5) for r in self.RATES))}

    def counters(self):
        torch = self.torch
        c = torch.zeros(4, dtype=torch.int64, device="cuda")
        for rate in self.RATES:
            info, ok, iters = self.out[rate]
            okb = ok.bool()
            c[0] += ok.numel()
            c[1] += (~okb).sum()
            bad = ((info[okb] ^ self.info_ref[rate][okb][:, : info.shape[1]]) & self.prot_mask[rate]) != 0
            c[2] += bad.any(dim=1).sum()
            c[3] += iters.sum()
        return c

    def counter_dict(self, c):
        return {"codewords": int(c[0]), "cw_fail": int(c[1]), "ok_but_wrong_protected_bits": int(c[2]),
                "mean_iters": float(c[3]) / max(1, int(c[0]))}

    def setup_e2e(self, n_e2e):
        self.e2e_n = n_e2e
        self.e2e_llr, self._pinned = {}, []
        for rate in self.RATES:
            host, _ = self.make_llr_host(rate, n_e2e, 555 + rate)
            pin = self.torch.empty(host.shape, dtype=self.torch.float32, pin_memory=True)
            pin.numpy()[:] = host
            self._pinned.append(pin)
            self.e2e_llr[rate] = pin.numpy()

    def step_e2e(self):
        return {rate: self.dec[rate].decode_batch_host(self.e2e_llr[rate]) for rate in self.RATES}

    def e2e_units(self):
        return len(self.RATES) * self.e2e_n / self.CW_PER_FRAME

    def e2e_bytes(self):
        from ria_b200 import fec
        h2d = len(self.RATES) * self.e2e_n * 648 * 4
        d2h = sum(self.e2e_n * ((fec.code_params(r)[0] + 7) // 8 + 1 + 4) for r in self.RATES)
        return h2d, d2h


def _cpu_ldpc_worker(args):
    n, seed, kind = args
    from oracle.bindings import Port, Ref
    wl = LdpcWorkload(n)
    impl = Ref() if kind == "reference" else Port()
    elapsed = 0.0
    for rate in wl.RATES:
        llr = wl.make_llr_host(rate, n, seed + rate)[0]
        t0 = time.perf_counter()
        impl.ldpc_decode_batch(rate, llr, wl.MAX_ITER[rate], wl.FACTOR)
        elapsed += time.perf_counter() - t0
    return elapsed


def cpu_baseline_ldpc(wl, per_rate_per_core):
    import multiprocessing as mp
    from oracle.bindings import Ref
    kind = "reference" if Ref.available() else "port"
    cores = host_cores()
    jobs = [(per_rate_per_core, 1000 + 17 * c, kind) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        busy = pool.map(_cpu_ldpc_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    per_core_cw = len(wl.RATES) * per_rate_per_core
    rate_sum = sum(per_core_cw / b for b in busy)
    return {"value": rate_sum / wl.CW_PER_FRAME, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{per_rate_per_core} codewords per rate per core x {cores} cores "
                      f"({cores * per_core_cw} codewords), {wall:.1f} s wall, {max(busy):.1f} s max busy"}


def measured_traffic(workload, kernel_name, frames_per_launch):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum from the
    committed ncu --set full capture (profiles/r1_traffic.json, bytes per frame) x frames per launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)
        return t[workload][kernel_name]["dram_bytes_per_frame"] * frames_per_launch
    except (OSError, KeyError, ValueError):
        return None


def make_workload(args):
    if args.workload == "mcdpsk":
        return McdpskC3Workload(args.batch or 100_000), cpu_baseline_mcdpsk, args.cpu_sample or 16
    if args.workload == "ldpc":
        return LdpcWorkload(args.batch or (1 << 20)), cpu_baseline_ldpc, args.cpu_sample or 4096
    return OfdmQam64Workload(args.batch or (1 << 20)), cpu_baseline_ofdm, args.cpu_sample or 6000


# ---------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------

def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, on
    a bounded sample of the same workload.  Rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import ria_b200  # noqa: F401  (host-side input synthesis only; no GPU work on this arm)
    wl, cpu_fn, sample = make_workload(args)
    if hasattr(wl, "pool_host"):
        wl.pool_host()                      # built once here, inherited by the forked workers
    for _ in range(args.warmup):
        cpu_fn(wl, max(16, sample // 16))
    runs = [cpu_fn(wl, sample) for _ in range(max(1, args.steps))]
    v = float(np.mean([r["value"] for r in runs]))
    last = runs[-1]
    units = last["cores"] * sample * (len(wl.RATES) / wl.CW_PER_FRAME if isinstance(wl, LdpcWorkload) else 1)
    line = {
        "impl": "reference", "metric": wl.metric, "value": v, "unit": wl.unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * units / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.describe(),
        "cpu_baseline": {"value": v, "unit": wl.unit, "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": v, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ria_b200", choices=["ria_b200", "reference"])
    ap.add_argument("--workload", default="ofdm_qam64", choices=["ofdm_qam64", "ldpc", "mcdpsk"])
    ap.add_argument("--batch", type=int, default=0, help="frames (or codewords per rate) per GPU")
    ap.add_argument("--e2e-batch", type=int, default=1 << 16)
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames (codewords per rate) per core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import ria_b200

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (ria_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    ctx = ria_b200.Context(local)
    stream = torch.cuda.Stream(device)
    wl, cpu_fn, cpu_sample = make_workload(args)
    warmup = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        wl.setup(ctx, device, rank, world)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(warmup):
            wl.step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ctx.set_timing(True)
        l0 = ctx.launch_count
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            wl.step()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = ctx.launch_count - l0
        kern_ms = {k: ctx.get_timing(k) for k in wl.kernels()}
        ctx.set_timing(False)
        clocks = sampler.stop() if rank == 0 else None

        # error counters: the only cross-GPU exchange of the path (one NCCL all-reduce)
        cnt = wl.counters()
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item()) / args.steps
        value = wl.units_per_step() * world / (ms_per_step * 1e-3)

        # end to end through the host-buffer C ABI (pinned host input, H2D + kernels + D2H timed)
        wl.setup_e2e(args.e2e_batch)
        wl.step_e2e()
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.step_e2e()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = wl.e2e_units() * world / float(te.item())
        h2d, d2h = wl.e2e_bytes()

    if rank == 0:
        peak, peak_src = measured_peaks()
        kinfo = wl.kernels()
        dom_kind = max(kern_ms, key=lambda k: kern_ms[k][0])          # most device time in the step
        dom_name, dom_bytes_step = kinfo[dom_kind]
        dom_ms, dom_n = kern_ms[dom_kind]
        kern_s = dom_ms / max(1, dom_n) * 1e-3
        bytes_per_launch = dom_bytes_step * args.steps / max(1, dom_n)
        achieved = bytes_per_launch / kern_s / 1e9
        traffic = measured_traffic(args.workload, dom_name, wl.units_per_step() * args.steps / max(1, dom_n))
        step_ms_local = ms / args.steps
        line = {
            "metric": wl.metric, "value": value, "unit": wl.unit,
            "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.describe(),
            "rx_msamples_per_s": wl.samples_per_step() * world / (ms_per_step * 1e-3) / 1e6,
            "counters": wl.counter_dict(cnt.cpu().numpy()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": ("profiles/r1_traffic.json: dram bytes per frame of this kernel from one "
                                            "ncu --set full capture x frames per launch" if traffic else None),
                         "peak_source": peak_src,
                         "kernel": dom_name, "kernel_ms_per_launch": kern_s * 1e3,
                         "launches_per_step": dom_n / args.steps,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "note": "dominant kernel = most device time in the step; see DESIGN.md section 3 "
                                 "for the per-frame byte counts and what bounds each kernel"},
            "kernels": {kinfo[k][0]: {"share_of_step": kern_ms[k][0] / args.steps / step_ms_local,
                                      "ms_per_step": kern_ms[k][0] / args.steps,
                                      "launches_per_step": kern_ms[k][1] / args.steps,
                                      "algorithmic_gb_per_s": (kinfo[k][1] * args.steps / (kern_ms[k][0] * 1e-3) / 1e9
                                                               if kern_ms[k][0] > 0 else None)}
                        for k in kern_ms},
            "e2e": {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "batch": getattr(wl, "e2e_n", args.e2e_batch),
                    "api": "ria_ofdm_rx_frames_host / ria_mcdpsk_rx_frames_host / ria_ldpc_decode_batch_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if hasattr(wl, "extra"):
            line.update(wl.extra(value))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_fn(wl, cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
