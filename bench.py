#!/usr/bin/env python
"""bench.py -- the measurement contract (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one pass of the receive hot path over one batch of synthetic received frames that
are already resident in HBM.  Rank 0 prints ONE JSON line.  For N > 1 the driver launches this
file under torch.distributed.run (one rank per GPU, NCCL); frames are sharded across ranks (weak
scaling: fixed per-GPU batch, no data-path collective) and the only exchange is one all-reduce of
the error counters.

Workloads
  ofdm_qam64      (default) BASELINE.json configs[3]: OFDM QAM64 R3/4 coherent, 15 pilots, 1024-FFT
                  CP 96, AWGN 28 dB, 1M frames per GPU: presynced demod -> de-interleave -> 4 x LDPC
                  -> header/CRC, i.e. "decoded frames/s".
  ofdm_qam64_cfo  the same chain on frames that arrive with the CFO / mixer phase of their sync
                  (integer TX carrier offsets in [-3, 3] Hz, told to the receiver with a +-0.2 Hz estimation
                  error): what production hands to the demodulator.  (SURVEY asks for U(-5, 5):
                  beyond +-3 Hz the REFERENCE demodulator's own LTS noise estimate drops below 29 dB even with
                  a perfect estimate and its QAM64 R3/4 frames stop decoding -- measured, DESIGN.md section 4 --
                  so wider offsets would time the retry ladder, not the CFO path.)
  ldpc            BASELINE.json configs[1]: LDPC R1/4..R3/4 decode of 1M codewords per rate.
  mcdpsk          BASELINE.json configs[2]: chirp-acquired MC-DPSK DBPSK 4x at -8 dB, HARQ chase combining.

The default run times ofdm_qam64 as the headline and then a short pass of each of the other three; their
results (with their own roofline / cpu_baseline / e2e) are in the line's "extra_workloads" object, so that
they are driver-measured too.  The CPU legs (cpu_baseline, --impl reference) live in oracle/cpu_bench.py and
never load the product library.
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
KK_OFDM_SYNC = 10
KK_ZC_SYNC = 5


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def sm_peaks():
    """SM-side denominators (fp32 rate, shared-memory bandwidth) for the kernels SURVEY.md 8(d) bounds by SM
    throughput: measured on this pool's B200 with profiles/peaks.cu (committed result), else nominal."""
    p = os.path.join(ROOT, "profiles", "r2_sm_peaks.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"fp32_tflops": float(d["fp32_fma_tflops"]), "fp32_unfused_tflops": float(d["fp32_unfused_tflops"]),
                "smem_gbs": 1e3 * float(d["smem_lds128_tb_per_s"]),
                "source": "measured (profiles/r2_sm_peaks.json: profiles/peaks.cu on this pool's B200)"}
    return {"fp32_tflops": 148 * 128 * 2 * 1.965e-3, "fp32_unfused_tflops": 148 * 128 * 1.965e-3,
            "smem_gbs": 148 * 128 * 1.965, "source": "nominal (148 SMs x 128 lanes / 128 B per clock x 1965 MHz)"}


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


def pin_rank_to_cores(local_rank: int, world: int):
    """Give every rank its own block of host cores (and, through first touch, host memory near them) so that
    eight ranks staging ~50 GB/s each do not pile onto one socket's cores.  No effect at world == 1."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return [mine[0], mine[-1]]
    except OSError:
        return None


def hbm_roofline(kern_name, kern_ms, n_launch, alg_bytes_step, steps, traffic=None, traffic_src=None):
    peak, peak_src = measured_peaks()
    kern_s = kern_ms / max(1, n_launch) * 1e-3
    bytes_per_launch = alg_bytes_step * steps / max(1, n_launch)
    achieved = bytes_per_launch / kern_s / 1e9 if kern_s > 0 else 0.0
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kern_name,
            "kernel_ms_per_launch": kern_s * 1e3, "launches_per_step": n_launch / steps,
            "algorithmic_bytes_per_launch": bytes_per_launch}


# ---------------------------------------------------------------------------------------------
# workload: OFDM QAM64 R3/4 full receive chain (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------

class OfdmQam64Workload:
    metric = "decoded_frames_per_s"
    unit = "frames/s"
    dtype = "f32"
    MOD, RATE, SPACING, SNR_DB = 8, 4, 4, 28.0          # QAM64, R3/4, pilot every 4th carrier
    FRAME_LEN, SYMBOL, PILOTS, DATA_CARRIERS = 12 * 1120, 1120, 15, 44
    CPU_SAMPLE = 6000

    def __init__(self, n_frames: int, cfo_span: float = 0.0):
        self.n = n_frames
        self.cfo_span = float(cfo_span)
        self.key = "ofdm_qam64_cfo" if cfo_span else "ofdm_qam64"
        self.name = "ofdm_qam64_r34_15pilots_awgn28dB" + (f"_synccfo{int(cfo_span)}Hz" if cfo_span else "")
        self.frame_len = self.FRAME_LEN

    def cpu_params(self):
        return {"modulation": self.MOD, "pilot_spacing": self.SPACING, "rate": self.RATE, "snr_db": self.SNR_DB,
                "pool": 64, "cfo_span": self.cfo_span}

    def cpu_baseline(self, sample):
        from oracle import cpu_bench
        return cpu_bench.cpu_ofdm(self.cpu_params(), sample)

    def describe(self):
        d = {"workload": self.name, "frames_per_gpu": self.n, "frame_samples": self.frame_len,
             "modulation": "QAM64", "code_rate": "R3/4", "fft": 1024, "cp": self.SYMBOL - 1024,
             "carriers": 59, "pilots": self.PILOTS, "data_carriers": self.DATA_CARRIERS,
             "channel": f"AWGN {self.SNR_DB} dB generated on the device (Philox); every frame has its own payload, "
                        "encodeFixedFrame + OFDM TX on the device (sample-identical to the reference transmitter)",
             "chain": "mix+CFO+CP+FFT1024+LTS/pilot est+MMSE+QAM64 LLR -> frame/channel "
                      "de-interleave -> complete v2::decodeFixedFrame (4x LDPC R3/4 0.9375/60 it, retry ladder and "
                      "false-positive repair armed: RIA_DECODE_FULL) -> header+CRC16",
             "l2": f"input batch ({self.n * self.frame_len * 4 / 1e9:.1f} GB) exceeds the 126 MB L2; no flush needed"}
        if self.cfo_span:
            d["cfo"] = (f"TX carrier offset by an integer in [-{self.cfo_span:.0f}, {self.cfo_span:.0f}] Hz per frame; the "
                        "receiver is told that offset +- U(0, 0.2) Hz (a sync estimate) and the mixer phase 0")
        return d

    def setup(self, ctx, device, rank, world):
        import torch
        import ria_b200
        from ria_b200 import ofdm, sim, txsynth
        self.torch, self.ctx = torch, ctx
        cfg = ofdm.ModemConfig.high_throughput(self.MOD)
        assert (cfg.getSymbolDuration(), cfg.getPilotCarriers(), cfg.getDataCarriers()) == (self.SYMBOL, self.PILOTS, self.DATA_CARRIERS)
        # Every frame is its own transmission: payload bytes (host RNG) -> v2 data frame (header, CRCs) ->
        # encodeFixedFrame, OFDM training + modulate and the AWGN channel on the device
        # (ria_encode_fixed_frame_batch_dev / ria_ofdm_tx_frames_dev are byte- / sample-identical to the
        # reference transmitter, tests/test_tx_gpu.py).  Generated in chunks so that only one chunk of clean
        # TX samples exists next to the resident RX batch.
        bpc = 60
        self.first_id = rank * self.n                       # global frame ids: sharding-independent payloads and noise
        bps = cfg.getDataCarriers() * 6
        self.samples = torch.empty((self.n, self.frame_len), dtype=torch.float32, device=device)
        self.sent_len = 4 * bpc - 2
        self.sent_dev = torch.empty((self.n, self.sent_len), dtype=torch.uint8, device=device)
        self.cfo = self.phase = None
        if self.cfo_span:
            g = torch.Generator(device=device).manual_seed(77 + rank)
            span = int(self.cfo_span)
            self.tx_off = torch.randint(-span, span + 1, (self.n,), device=device, generator=g)
            self.cfo = (self.tx_off.float() + (torch.rand(self.n, device=device, generator=g) * 0.4 - 0.2)).contiguous()
            self.phase = torch.zeros(self.n, dtype=torch.float32, device=device)
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
            gid = self.first_id + off
            if self.cfo_span:
                # a transmitter whose carrier is off by d Hz IS the frequency-shifted signal: re-modulate the
                # frames of each offset group with center_freq + d (row i here is frame i of the chunk; the
                # roll below only re-orders rows for the channel's pool indexing)
                offs = self.tx_off[off:off + m]
                for d in range(-int(self.cfo_span), int(self.cfo_span) + 1):
                    idx = (offs == d).nonzero().flatten()
                    if d == 0 or idx.numel() == 0:
                        continue
                    cfg_d = ofdm.ModemConfig.from_buffer_copy(bytes(cfg))
                    cfg_d.center_freq = cfg.center_freq + d
                    tx[idx] = ofdm.ofdm_tx_frames(cfg_d, coded[idx].contiguous(), ctx)
            # the channel takes TX row (global id % pool rows) and keys its noise by the global id: with the
            # chunk as the pool, rotate it so that frame i of the chunk sits in row (gid + i) % m
            if gid % m:
                tx = torch.roll(tx, shifts=gid % m, dims=0)
            sim.awgn_batch(tx, m, self.SNR_DB, seed=2026, first_frame_id=gid, out=self.samples[off:off + m], ctx=ctx)
            del tx, coded, fr_dev
        ctx.set_decode_flags(ria_b200.DECODE_FULL)          # the reference's complete decodeFixedFrame
        self.chain = ofdm.OfdmRxChain(cfg, self.RATE, True, ctx)
        self.out = None
        torch.cuda.synchronize()

    def release(self):
        self.samples = self.sent_dev = self.out = self.cfo = self.phase = None
        self._pin = self.e2e_host = None

    def step(self):
        self.out = self.chain.process_batch(self.samples, self.cfo, self.phase)

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
            KK_OFDM_PHASE: ("ofdm_phase_scan_kernel (CFO phase accumulator, CFO frames only)", self.n * 8 if self.cfo_span else 0),
            KK_OFDM_FFT: ("ofdm_fft_kernel", self.n * (samples + bins)),
            KK_OFDM_CARRIER: ("ofdm_carrier2_kernel", self.n * (bins + llr)),
            KK_OFDM_DEMOD: ("ofdm_presynced_kernel (residual-CFO re-run frames only)", 0),
            KK_LDPC: ("ldpc_decode_kernel", self.n * (llr + 4 * (61 + 5))),
            KK_FRAME_STATUS: ("frame_status_kernel", self.n * (4 * 72 + 4 * 5 + 40)),
            KK_LDPC_RETRY: ("ldpc_fail_list_kernel + ldpc_retry_kernel (retry ladder, failed frames only)", self.n * 4),
            KK_FRAME_REPAIR: ("frame_repair_list_kernel + frame_repair_kernel (false-positive repair, invalid frames only)",
                              self.n * (4 + 4 * 60)),
        }

    def roofline(self, kern_ms, steps, counters):
        """dominant kernel = most device time in the step; HBM bound (demod / FFT / demap stages)"""
        kinfo = self.kernels()
        dom = max(kern_ms, key=lambda k: kern_ms[k][0])
        name, bytes_step = kinfo[dom]
        ms, n_launch = kern_ms[dom]
        traffic = measured_traffic("ofdm_qam64", name, self.n * steps / max(1, n_launch))
        r = hbm_roofline(name, ms, n_launch, bytes_step, steps, traffic,
                         "profiles/r*_traffic.json: dram bytes per frame of this kernel from one ncu --set full capture "
                         "x frames per launch" if traffic else None)
        r["note"] = "dominant kernel = most device time in the step; see DESIGN.md section 3 for the per-frame byte counts"
        return r

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
        self.e2e_cfo = self.cfo[:n_e2e].cpu().numpy() if self.cfo is not None else None
        self.e2e_phase = self.phase[:n_e2e].cpu().numpy() if self.phase is not None else None

    def step_e2e(self):
        return self.chain.process_batch_host(self.e2e_host, self.e2e_cfo, self.e2e_phase)

    def e2e_units(self):
        return float(self.e2e_n)

    def e2e_bytes(self):
        from ria_b200 import ofdm
        # ria_ofdm_rx_frames_host copies only the 1024-sample FFT window of each 1120-sample symbol (the
        # cyclic prefix is never read by the demodulator): 12 x 1024 x 4 B per frame cross PCIe
        sym_copied = 1024 if os.environ.get("RIA_H2D_FULL_SYMBOLS") != "1" else 1120
        return (self.e2e_n * (self.frame_len // 1120) * sym_copied * 4 + (8 * self.e2e_n if self.cfo_span else 0),
                self.e2e_n * (240 + ofdm.FRAME_STATUS_DTYPE.itemsize + 4))


# ---------------------------------------------------------------------------------------------
# workload: OFDM_COX -- the same frames behind a Schmidl-Cox preamble, acquired by the batched searchForSync
# (SURVEY.md 8f rank 4): window -> search -> frames at the LTS position with the CFO found -> the chain above
# ---------------------------------------------------------------------------------------------
class OfdmCoxWorkload(OfdmQam64Workload):
    LEAD, WINDOW = 2000, 24000                          # [2000 quiet][silent symbol][4 STS][2 LTS][10 data symbols][tail]
    CPU_SAMPLE = 150
    e2e_api = "ria_ofdm_cox_rx_frames_host (pinned host windows)"

    def __init__(self, n_frames: int):
        super().__init__(n_frames)
        self.key = "ofdm_cox"
        self.name = "ofdm_cox_qam64_r34_15pilots_awgn28dB_schmidl_cox_acquisition"

    def cpu_params(self):
        p = super().cpu_params()
        p["cox"] = {"lead": self.LEAD, "window": self.WINDOW, "frame_len": self.FRAME_LEN}
        p["pool"] = 16
        return p

    def describe(self):
        d = super().describe()
        d["window_samples"] = self.WINDOW
        d["chain"] = ("ria_ofdm_cox_rx_frames_dev: searchForSync (energy walk, Schmidl-Cox metric, plateau, LTS timing, coarse "
                      "CFO) -> frame at the LTS position with the CFO / phase found -> " + d["chain"])
        d["l2"] = f"input batch ({self.n * self.WINDOW * 4 / 1e9:.1f} GB) exceeds the 126 MB L2; no flush needed"
        return d

    def setup(self, ctx, device, rank, world):
        import torch
        import ria_b200
        from ria_b200 import ofdm, sim, sync, txsynth
        self.torch, self.ctx, self.sync = torch, ctx, sync
        cfg = self.cfg = ofdm.ModemConfig.high_throughput(self.MOD)
        bpc = 60
        self.first_id = rank * self.n
        bps = cfg.getDataCarriers() * 6
        self.windows = torch.empty((self.n, self.WINDOW), dtype=torch.float32, device=device)
        self.sent_len = 4 * bpc - 2
        self.sent_dev = torch.empty((self.n, self.sent_len), dtype=torch.uint8, device=device)
        chunk = 32768
        for off in range(0, self.n, chunk):
            m = min(chunk, self.n - off)
            rng = np.random.default_rng([2027, self.first_id + off])
            frames = txsynth.make_data_frames("K1ABC", "W2XYZ", self.first_id + off,
                                              rng.integers(0, 256, size=(m, 4 * bpc - 19 - 2), dtype=np.uint8), bpc)
            fr_dev = torch.from_numpy(frames).to(device)
            self.sent_dev[off:off + m] = fr_dev
            coded = ofdm.encode_fixed_frame_batch(fr_dev, self.RATE, True, bps, ctx)
            tx = ofdm.ofdm_cox_tx_frames(cfg, coded, ctx)
            rows = torch.zeros((m, self.WINDOW), dtype=torch.float32, device=device)
            rows[:, self.LEAD:self.LEAD + tx.shape[1]] = tx
            gid = self.first_id + off
            if gid % m:
                rows = torch.roll(rows, shifts=gid % m, dims=0)
            sim.awgn_batch(rows, m, self.SNR_DB, seed=2027, first_frame_id=gid, out=self.windows[off:off + m], ctx=ctx)
            del tx, coded, fr_dev, rows
        ctx.set_decode_flags(ria_b200.DECODE_FULL)
        self.chain = ofdm.OfdmCoxRxChain(cfg, self.RATE, True, ctx)
        self.out = None
        torch.cuda.synchronize()

    def release(self):
        self.windows = self.sent_dev = self.out = None
        self._pin = self.e2e_host = None

    def _run(self, windows):
        torch = self.torch
        data, status, snr, sync = self.chain.process_windows(windows, self.FRAME_LEN, 0.8)
        self.detected = sync.view(torch.int32)[:, 0] != 0
        return data, status, snr

    def step(self):
        self.out = self._run(self.windows)

    def samples_per_step(self):
        return float(self.n) * self.WINDOW

    def kernels(self):
        k = super().kernels()
        k[KK_OFDM_PHASE] = ("ofdm_phase_scan_kernel (CFO phase accumulator)", self.n * 8)
        k[KK_OFDM_SYNC] = ("cox_scan_kernel + cox_decide_kernel (Schmidl-Cox acquisition)", self.n * (self.WINDOW * 4 + 32))
        return k

    # per window of this layout: the LTS passband search at 3921 offsets x 1120 taps x (3 multiplies + 3 adds) = 26.3 Mflop,
    # plus ~74 Schmidl-Cox metrics (35 visited positions, 38 plateau points, the CFO estimate) of two 1024-point
    # transforms (2 x 5 N log2 N) and 4 x 512-term sums each = 8.1 Mflop
    COX_FLOP_PER_WINDOW = 3921 * 1120 * 6 + 74 * (2 * 5 * 1024 * 10 + 7 * 1024)

    def roofline(self, kern_ms, steps, counters):
        """the acquisition is arithmetic on shared-memory tiles: fp32 bound (un-fused multiplies and adds, the order of
        every sum being the reference's); HBM traffic is the 96 KB window"""
        sp = sm_peaks()
        ms, n_launch = kern_ms[KK_OFDM_SYNC]
        kern_s = ms / steps * 1e-3
        achieved = self.n * self.COX_FLOP_PER_WINDOW / kern_s / 1e12 if kern_s > 0 else 0.0
        hbm_peak, _ = measured_peaks()
        return {"bound": "fp32", "achieved": achieved, "peak": sp["fp32_tflops"], "unit": "TFLOP/s",
                "frac": achieved / sp["fp32_tflops"], "traffic": None, "peak_source": sp["source"],
                "kernel": self.kernels()[KK_OFDM_SYNC][0], "kernel_ms_per_launch": ms / max(1, n_launch),
                "launches_per_step": n_launch / steps, "algorithmic_flop_per_step": self.n * self.COX_FLOP_PER_WINDOW,
                "windows_per_s": self.n / kern_s if kern_s > 0 else None,
                "hbm_frac_on_algorithmic_bytes": (self.n * self.WINDOW * 4 / kern_s / 1e9 / hbm_peak) if kern_s > 0 else None,
                "note": "peak = measured FMA rate; the kernel may not fuse (bit-exact sums), so 0.5 is its ceiling"}

    def counters(self):
        c = super().counters()
        c = self.torch.cat([c, self.detected.sum().reshape(1).to(c.dtype)])
        return c

    def counter_dict(self, c):
        d = super().counter_dict(c)
        d["sync_found"] = int(c[5])
        return d

    def setup_e2e(self, n_e2e):
        torch = self.torch
        self.e2e_n = n_e2e = min(n_e2e, self.n)
        pin = torch.empty((n_e2e, self.WINDOW), dtype=torch.float32, pin_memory=True)
        pin.copy_(self.windows[:n_e2e])
        torch.cuda.synchronize()
        self._pin = pin
        self.e2e_host = pin.numpy()

    def step_e2e(self):
        return self.chain.process_windows_host(self.e2e_host, self.FRAME_LEN, 0.8)

    def e2e_bytes(self):
        from ria_b200 import ofdm
        return (self.e2e_n * self.WINDOW * 4, self.e2e_n * (240 + ofdm.FRAME_STATUS_DTYPE.itemsize + 4))


# ---------------------------------------------------------------------------------------------
# workload: chirp-acquired MC-DPSK 4x spread with HARQ chase combining (BASELINE.json configs[2])
# ---------------------------------------------------------------------------------------------

class McdpskC3Workload:
    """Every frame is received twice (HARQ): per reception the channel is simulated on the device
    (AWGN at -8 dB over the whole transmission), the dual chirp is searched in the first 120 000
    samples of the row (streaming_decoder.cpp:408-411), the frame is demodulated at the detected
    training start with the detected CFO, its soft bits are chase-combined with the previous
    reception and the LDPC R1/4 codeword is decoded.  A step is both receptions of all frames."""
    key = "mcdpsk"
    name = "mcdpsk_dbpsk_10car_4x_awgn-8dB_chirp_chase"
    metric = "decoded_frames_per_s"
    unit = "frames/s"
    dtype = "f32"
    RATE, MAX_ITER, FACTOR, SNR_DB, POOL, LEAD, TAIL, WINDOW = 0, 50, 0.9375, -8.0, 64, 2000, 800, 120000
    PREAMBLE, FRAME_LEN = 57600, 4096 + 512 + 133120         # dual chirp; 8 training + 1 reference + 65 x 4 data symbols
    CHIRP_FLOP_PER_WINDOW = 4 * 5 * 131072 * 17              # two (FFT + IFFT) pairs of 131072 points
    CPU_SAMPLE = 16
    CHIRP_RECEPTIONS = 2

    def __init__(self, n_frames: int):
        self.n = n_frames
        self.frame_len = self.FRAME_LEN
        self.row_len = self.LEAD + self.PREAMBLE + self.FRAME_LEN + self.TAIL

    def cpu_params(self):
        return {"pool": 16, "lead": self.LEAD, "tail": self.TAIL, "rate": self.RATE, "snr_db": self.SNR_DB,
                "window": self.WINDOW, "max_iter": self.MAX_ITER, "factor": self.FACTOR}

    def cpu_baseline(self, sample):
        from oracle import cpu_bench
        return cpu_bench.cpu_mcdpsk(self.cpu_params(), sample)

    def cfg(self):
        from ria_b200 import mcdpsk
        return mcdpsk.MultiCarrierDPSKConfig.level4_dbpsk(mcdpsk.SPREAD_4X)

    def describe(self):
        return {"workload": self.name, "frames_per_gpu": self.n, "receptions_per_frame": 2,
                "row_samples": self.row_len, "frame_samples": self.frame_len, "modulation": "DBPSK",
                "carriers": 10, "spreading": 4, "code_rate": "R1/4",
                "channel": f"AWGN {self.SNR_DB} dB over the whole transmission, generated on the device "
                           f"(Philox) inside the timed step, {self.POOL} distinct TX frames",
                "chain": "per reception: AWGN -> dual-chirp sync (131072-pt FFT matched filter) -> Hilbert CFO "
                         "correction -> 10-carrier correlation demod + 4x despreading -> chase combine -> "
                         "LDPC R1/4 (0.9375, 50 it)",
                "l2": f"one reception of the batch ({self.n * self.row_len * 4 / 1e9:.1f} GB) exceeds the 126 MB L2; no flush needed"}

    def setup(self, ctx, device, rank, world):
        import torch
        from ria_b200 import mcdpsk, sync, txsynth
        self.torch, self.ctx, self.device = torch, ctx, device
        # TX pool on the device: LDPC-encoded payloads -> MC-DPSK modulator, behind the dual-chirp preamble
        rng = np.random.default_rng(31)
        data = rng.integers(0, 256, size=(self.POOL, 20), dtype=np.uint8)
        info_bits = np.concatenate([np.unpackbits(data, axis=1), np.zeros((self.POOL, 2), np.uint8)], axis=1)    # k = 162
        cw = np.packbits(txsynth.ldpc_encode_bits(info_bits, self.RATE), axis=1)
        body = mcdpsk.mcdpsk_tx_frames(self.cfg(), torch.from_numpy(cw).to(device), ctx)
        pre = sync.chirp_generate(device=device, ctx=ctx)
        assert body.shape[1] == self.FRAME_LEN and pre.numel() == self.PREAMBLE, (body.shape, pre.numel())
        pool = torch.zeros((self.POOL, self.row_len), dtype=torch.float32, device=device)
        pool[:, self.LEAD:self.LEAD + self.PREAMBLE] = pre
        pool[:, self.LEAD + self.PREAMBLE:self.LEAD + self.PREAMBLE + self.FRAME_LEN] = body
        self.pool_dev = pool
        self.sent_dev = torch.from_numpy(data).to(device)
        self.first_id = rank * self.n
        self.rows = torch.empty((self.n, self.row_len), dtype=torch.float32, device=device)
        self.acc = torch.empty((self.n, 648), dtype=torch.float32, device=device)
        self.chain = mcdpsk.McdpskRxChain(self.cfg(), self.RATE, self.MAX_ITER, self.FACTOR, 0.15, ctx)
        self.out = None
        self.first_ok = None
        self.epoch = 0
        torch.cuda.synchronize()

    def release(self):
        self.rows = self.acc = self.out = self.pool_dev = self._pin = None

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
            KK_CHIRP_SYNC: ("chirp_* (131072-pt matched filter for both chirps + peak search)", n2 * win),
            KK_MCDPSK_CFO: ("mcdpsk_phase_scan_kernel + mcdpsk_cfo_kernel", n2 * 2 * body),
            KK_MCDPSK: ("mcdpsk_demod_kernel", n2 * (body + 652 * 4)),
            KK_CHASE: ("chase_combine_kernel", n2 * 3 * 648 * 4),
            KK_LDPC: ("ldpc_decode_kernel", n2 * (648 * 4 + 24 + 5)),
        }

    def roofline(self, kern_ms, steps, counters):
        """The matched filter dominates; SURVEY.md 8(d) bounds it by fp32 issue (44.6 Mflop per window for the
        reference's two FFT + IFFT pairs), HBM traffic being the 480 KB window when staged once."""
        sp = sm_peaks()
        ms, n_launch = kern_ms[KK_CHIRP_SYNC]
        windows_per_launch = float(self.CHIRP_RECEPTIONS) * self.n * steps / max(1, n_launch)
        kern_s = ms / max(1, n_launch) * 1e-3
        achieved = windows_per_launch * self.CHIRP_FLOP_PER_WINDOW / kern_s / 1e12 if kern_s > 0 else 0.0
        hbm_peak, _ = measured_peaks()
        return {"bound": "fp32", "achieved": achieved, "peak": sp["fp32_tflops"], "unit": "TFLOP/s",
                "frac": achieved / sp["fp32_tflops"], "traffic": chirp_traffic(windows_per_launch),
                "peak_source": sp["source"], "kernel": self.kernels()[KK_CHIRP_SYNC][0],
                "kernel_ms_per_launch": kern_s * 1e3, "launches_per_step": n_launch / steps,
                "algorithmic_flop_per_launch": windows_per_launch * self.CHIRP_FLOP_PER_WINDOW,
                "windows_per_s": windows_per_launch / kern_s if kern_s > 0 else None,
                "hbm_frac_on_algorithmic_bytes": (windows_per_launch * self.WINDOW * 4 / kern_s / 1e9 / hbm_peak) if kern_s > 0 else None,
                "note": "algorithmic flops = SURVEY.md 8(d) K2 figure (4 x 5 N log2 N, N = 131072) per window; the share of the "
                        "step is in `kernels`"}

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


# ---------------------------------------------------------------------------------------------
# workload: configs[2] with the retransmission received the way a connected station receives it: first reception
# behind the dual chirp, HARQ retransmission behind the Zadoff-Chu data preamble ("ZC + CHIRP acquisition")
# ---------------------------------------------------------------------------------------------
class McdpskZcRetxWorkload(McdpskC3Workload):
    key = "mcdpsk_zc_retx"
    name = "mcdpsk_dbpsk_10car_4x_awgn-8dB_chirp_then_zc_retransmission_chase"
    ZC_WINDOW = 31120                                    # connected-mode search window (streaming_decoder.cpp:423-435)
    CHIRP_RECEPTIONS = 1

    def kernels(self):
        k = super().kernels()
        n = self.n
        k[KK_CHIRP_SYNC] = (k[KK_CHIRP_SYNC][0], n * self.WINDOW * 4)
        k[KK_ZC_SYNC] = ("zc_baseband_kernel + zc_coarse_tile_kernel + zc_finish_kernel", n * self.ZC_WINDOW * 4)
        return k

    def describe(self):
        d = super().describe()
        d["chain"] = ("reception 1: " + d["chain"].replace("per reception: ", "") + "; reception 2 (HARQ retransmission, connected "
                      "mode): AWGN -> Zadoff-Chu data-preamble sync (roots DATA | CONTROL, 31 120-sample window) -> same demod "
                      "-> chase combine with reception 1 -> LDPC")
        return d

    def setup(self, ctx, device, rank, world):
        super().setup(ctx, device, rank, world)
        import torch
        from ria_b200 import sync
        pre = sync.zc_preamble(root=2, device=device, ctx=ctx)       # DATA root
        body = self.pool_dev[:, self.LEAD + self.PREAMBLE:self.LEAD + self.PREAMBLE + self.FRAME_LEN]
        self.row_len2 = self.LEAD + pre.numel() + self.FRAME_LEN + self.TAIL
        pool2 = torch.zeros((self.POOL, self.row_len2), dtype=torch.float32, device=device)
        pool2[:, self.LEAD:self.LEAD + pre.numel()] = pre
        pool2[:, self.LEAD + pre.numel():self.LEAD + pre.numel() + self.FRAME_LEN] = body
        self.pool2_dev = pool2
        self.rows2 = self.rows.view(-1)[: self.n * self.row_len2].view(self.n, self.row_len2)    # same storage, narrower rows
        torch.cuda.synchronize()

    def release(self):
        self.pool2_dev = self.rows2 = None
        super().release()

    def step(self):
        from ria_b200 import sim
        sim.awgn_batch(self.pool_dev, self.n, self.SNR_DB, seed=9000 + 2 * self.epoch, first_frame_id=self.first_id,
                       out=self.rows, ctx=self.ctx)
        self.out = self.chain.process_batch(self.rows, self.frame_len, self.WINDOW, self.acc, True, self.out)
        self.first_ok = self.out["ok"].clone()
        sim.awgn_batch(self.pool2_dev, self.n, self.SNR_DB, seed=9001 + 2 * self.epoch, first_frame_id=self.first_id,
                       out=self.rows2, ctx=self.ctx)
        self.out = self.chain.process_batch_zc(self.rows2, self.frame_len, self.ZC_WINDOW, self.acc, False, self.out)
        self.epoch += 1

    def samples_per_step(self):
        return float(self.n) * (self.row_len + self.row_len2)


# ---------------------------------------------------------------------------------------------
# workload: batched LDPC decode (BASELINE.json configs[1])
# ---------------------------------------------------------------------------------------------

class LdpcWorkload:
    """configs[1]: R1/4..R3/4 min-sum decode of 1M codewords per rate from synthetic AWGN LLRs
    (LLR model of tools/test_chase_cache.cpp:20-34; Es/N0 and decoder settings of SURVEY.md 8d)."""

    key = "ldpc"
    name = "ldpc_r14_r34_1M_cw_per_rate"
    metric = "decoded_frames_per_s"
    unit = "frames/s"
    dtype = "f32"
    RATES = (0, 2, 3, 4)
    ESN0 = {0: 1.0, 2: 4.0, 3: 6.0, 4: 7.0}
    MAX_ITER = {0: 50, 2: 80, 3: 70, 4: 60}
    EDGES = {0: 2437, 2: 1623, 3: 1510, 4: 1134}         # edges of the generated H per rate
    FACTOR = 0.9375
    CW_PER_FRAME = 4
    CPU_SAMPLE = 4096

    def __init__(self, n_cw: int):
        self.n_cw = n_cw

    def cpu_params(self):
        return {"rates": self.RATES, "esn0": self.ESN0, "max_iter": self.MAX_ITER, "factor": self.FACTOR,
                "cw_per_frame": self.CW_PER_FRAME}

    def cpu_baseline(self, sample):
        from oracle import cpu_bench
        return cpu_bench.cpu_ldpc(self.cpu_params(), sample)

    def describe(self):
        return {"workload": self.name, "codewords_per_rate_per_gpu": self.n_cw,
                "rates": "R1/4,R1/2,R2/3,R3/4", "esn0_db": "1,4,6,7", "min_sum_factor": self.FACTOR,
                "max_iter": "50,80,70,60", "frame": "4 codewords (v2 fixed frame)",
                "l2": f"inputs ({self.n_cw * 648 * 4 / 1e9:.1f} GB per rate) exceed the 126 MB L2; no flush needed"}

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
            k, _, e = fec.code_params(rate)
            assert e == self.EDGES[rate], (rate, e)
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

    def release(self):
        self.llr = self.out = self.info_ref = None
        self._pinned = self.e2e_llr = None

    def step(self):
        self.out = {rate: self.dec[rate].decode_batch(self.llr[rate]) for rate in self.RATES}

    def units_per_step(self):
        return len(self.RATES) * self.n_cw / self.CW_PER_FRAME

    def samples_per_step(self):
        return 0.0

    def kernels(self):
        from ria_b200 import fec
        return {KK_LDPC: ("ldpc_decode_kernel",
                          sum(self.n_cw * (648 * 4 + (fec.code_params(r)[0] + 7) // 8 + 5) for r in self.RATES))}

    def roofline(self, kern_ms, steps, counters):
        """shared-memory bound, messages live on chip.  Algorithmic shared-memory bytes =
        16 B per edge per pass (v2c read, c2v write, c2v read for the totals, parity term), passes = iterations
        + 1 per codeword; the count uses the decoder's own iteration counters."""
        sp = sm_peaks()
        ms, n_launch = kern_ms[KK_LDPC]
        smem_bytes = edge_updates = 0.0
        for rate in self.RATES:
            passes = float(self.out[rate][2].sum().item()) + self.n_cw          # sum(iters) + 1 per codeword
            smem_bytes += 16.0 * self.EDGES[rate] * passes
            edge_updates += 2.0 * self.EDGES[rate] * passes
        kern_s_step = ms / steps * 1e-3
        achieved = smem_bytes / kern_s_step / 1e9 if kern_s_step > 0 else 0.0
        hbm = hbm_roofline("ldpc_decode_kernel", ms, n_launch, self.kernels()[KK_LDPC][1], steps)
        return {"bound": "smem", "achieved": achieved, "peak": sp["smem_gbs"], "unit": "GB/s", "frac": achieved / sp["smem_gbs"],
                "traffic": None, "peak_source": sp["source"], "kernel": "ldpc_decode_kernel",
                "kernel_ms_per_launch": ms / max(1, n_launch), "launches_per_step": n_launch / steps,
                "algorithmic_smem_bytes_per_step": smem_bytes,
                "edge_updates_per_s": edge_updates / kern_s_step if kern_s_step > 0 else None,
                "hbm_frac_on_algorithmic_bytes": hbm["frac"],
                "note": "16 B x E x (iterations + 1) per codeword (SURVEY.md 8d K11) over the kernel's device time"}

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


def _traffic_file():
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return p
    return None


def measured_traffic(workload, kernel_name, frames_per_launch):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum from the
    committed ncu --set full capture (profiles/r*_traffic.json, bytes per frame) x frames per launch."""
    try:
        with open(_traffic_file()) as f:
            t = json.load(f)
        return t[workload][kernel_name]["dram_bytes_per_frame"] * frames_per_launch
    except (OSError, KeyError, ValueError, TypeError):
        return None


def chirp_traffic(windows_per_launch):
    try:
        with open(_traffic_file()) as f:
            t = json.load(f)
        return t["mcdpsk"]["chirp"]["dram_bytes_per_window"] * windows_per_launch
    except (OSError, KeyError, ValueError, TypeError):
        return None


def make_workload(name, batch):
    if name == "mcdpsk":
        return McdpskC3Workload(batch or 100_000)
    if name == "mcdpsk_zc_retx":
        return McdpskZcRetxWorkload(batch or 100_000)
    if name == "ldpc":
        return LdpcWorkload(batch or (1 << 20))
    if name == "ofdm_qam64_cfo":
        return OfdmQam64Workload(batch or (1 << 18), cfo_span=3.0)
    if name == "ofdm_cox":
        return OfdmCoxWorkload(batch or (1 << 16))
    return OfdmQam64Workload(batch or (1 << 20))


# ---------------------------------------------------------------------------------------------
# --impl reference
# ---------------------------------------------------------------------------------------------

def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, on
    a bounded sample of the same workload.  Rank 0 only.  Nothing on this arm imports ria_b200: the inputs
    come from the reference's own transmitter (oracle/cpu_bench.py)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    wl = make_workload(args.workload, args.batch)
    sample = args.cpu_sample or wl.CPU_SAMPLE
    for _ in range(args.warmup):
        wl.cpu_baseline(max(16, sample // 16))
    runs = [wl.cpu_baseline(sample) for _ in range(max(1, args.steps))]
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
        "product_library_loaded": "ria_b200" in sys.modules,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# one workload on this rank's GPU
# ---------------------------------------------------------------------------------------------

def run_workload(wl, ctx, stream, device, rank, world, local, steps, warmup, e2e_batch, cpu_sample, with_cpu):
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        wl.setup(ctx, device, rank, world)
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
        for _ in range(steps):
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
        ms_per_step = float(t.item()) / steps
        value = wl.units_per_step() * world / (ms_per_step * 1e-3)
        roofline = wl.roofline(kern_ms, steps, cnt) if rank == 0 else None

        # end to end through the host-buffer C ABI (pinned host input, H2D + kernels + D2H timed)
        wl.setup_e2e(e2e_batch)
        wl.step_e2e()
        barrier()
        e2e_steps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.step_e2e()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = wl.e2e_units() * world / float(te.item())
        h2d, d2h = wl.e2e_bytes()

    res = None
    if rank == 0:
        kinfo = wl.kernels()
        step_ms_local = ms / steps
        res = {
            "metric": wl.metric, "value": value, "unit": wl.unit,
            "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.describe(),
            "rx_msamples_per_s": wl.samples_per_step() * world / (ms_per_step * 1e-3) / 1e6,
            "counters": wl.counter_dict(cnt.cpu().numpy()),
            "roofline": roofline,
            "kernels": {kinfo[k][0]: {"share_of_step": kern_ms[k][0] / steps / step_ms_local,
                                      "ms_per_step": kern_ms[k][0] / steps,
                                      "launches_per_step": kern_ms[k][1] / steps,
                                      "algorithmic_gb_per_s": (kinfo[k][1] * steps / (kern_ms[k][0] * 1e-3) / 1e9
                                                               if kern_ms[k][0] > 0 else None)}
                        for k in kern_ms},
            "e2e": {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "batch": getattr(wl, "e2e_n", e2e_batch),
                    "api": getattr(wl, "e2e_api", "ria_ofdm_rx_frames_host / ria_mcdpsk_rx_frames_host / ria_ldpc_decode_batch_host "
                                                  "(pinned host buffers)")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if hasattr(wl, "extra"):
            res.update(wl.extra(value))
        if with_cpu:
            res["cpu_baseline"] = wl.cpu_baseline(cpu_sample or wl.CPU_SAMPLE)
    wl.release()
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ria_b200", choices=["ria_b200", "reference"])
    ap.add_argument("--workload", default="ofdm_qam64", choices=["ofdm_qam64", "ofdm_qam64_cfo", "ofdm_cox", "ldpc", "mcdpsk", "mcdpsk_zc_retx"])
    ap.add_argument("--batch", type=int, default=0, help="frames (or codewords per rate) per GPU")
    ap.add_argument("--e2e-batch", type=int, default=1 << 16)
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames (codewords per rate) per core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short passes of the other workloads")
    ap.add_argument("--no-cli-simulator", action="store_true", help="skip the cli_simulator-per-core CPU line")
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
    affinity = pin_rank_to_cores(local, world)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    ctx = ria_b200.Context(local)
    stream = torch.cuda.Stream(device)
    warmup = max(args.warmup, 3)
    with_cpu = world == 1 and not args.no_cpu_baseline
    line = run_workload(make_workload(args.workload, args.batch), ctx, stream, device, rank, world, local,
                        args.steps, warmup, args.e2e_batch, args.cpu_sample, with_cpu)
    if rank == 0 and affinity:
        line["rank_core_affinity"] = f"each rank pinned to its own block of host cores (rank 0: {affinity[0]}..{affinity[1]})"

    # the other BASELINE configs, a short pass each (driver-measured through this same line)
    if args.workload == "ofdm_qam64" and not args.batch and not args.no_extras:
        extras = {}
        short = max(2, min(args.steps, 4))
        for name, batch in (("ofdm_qam64_cfo", 1 << 18), ("ofdm_cox", 1 << 15), ("mcdpsk", 1 << 15), ("mcdpsk_zc_retx", 1 << 15),
                            ("ldpc", 1 << 20)):
            ctx.set_decode_flags(0)
            r = run_workload(make_workload(name, batch), ctx, stream, device, rank, world, local,
                             short, 3, min(args.e2e_batch, 1 << 15), 0, with_cpu)
            if rank == 0:
                extras[name] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "counters", "roofline",
                                                  "kernels", "e2e", "gpu_launches", "clocks", "cpu_baseline", "receptions_per_s",
                                                  "rx_msamples_per_s") if k in r}
        if rank == 0:
            line["extra_workloads"] = extras
            line["gpu_launches_headline"] = line["gpu_launches"]
            line["gpu_launches"] = int(line["gpu_launches"] + sum(e["gpu_launches"] for e in extras.values()))
            if with_cpu and not args.no_cli_simulator:
                from oracle import cpu_bench
                line["cpu_baseline"]["cli_simulator_per_core"] = cpu_bench.cli_simulator_per_core()

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
