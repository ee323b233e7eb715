#!/usr/bin/env python
"""Adaptive-waveform SNR sweep over a Watterson channel (BASELINE.json configs[4]).

For every SNR point the reference's selection ladder (protocol::recommendWaveformAndRate,
src/protocol/waveform_selection.hpp:112-222) picks waveform, modulation, code rate and spreading;
the transmissions (dual-chirp preamble + frame) are synthesised once per mode on the device, the HF
channel (sim::WattersonChannel, src/sim/hf_channel.hpp) runs on the device for every reception, every
reception goes through acquisition (dual-chirp sync), demodulation at the detected start with the
detected CFO and the decoder of that mode, and the frame-error counters are summed over the GPUs with
one NCCL all-reduce per SNR point.  The ladder is fed with the fading index the demodulator itself
reports on the channel (one re-selection per point).  Frames are sharded over ranks; global frame ids key
the channel's random streams, so the result does not depend on the number of GPUs.

    python sweep.py --frames 4096 --condition moderate
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 sweep.py --frames 65536 --condition poor

Rank 0 prints one JSON line per SNR point and a final summary line.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

CONDITIONS = {"awgn": 0, "good": 1, "moderate": 2, "poor": 3, "flutter": 4}
# fading index the ladder is fed with (what the demodulators report on such channels)
FADING_INDEX = {"awgn": 0.0, "good": 0.15, "moderate": 0.45, "poor": 0.9, "flutter": 0.9}
POOL = 16


LEAD, TAIL, CHUNK = 2000, 800, 16384


class ModeCache:
    """TX pools and receive chains per (waveform, modulation, rate, spreading).  A pool row is one whole transmission as
    a connecting station sends it: [silence][dual-chirp preamble][training ... data][silence]."""

    def __init__(self, ctx, device):
        self.ctx, self.device, self.modes = ctx, device, {}
        self._chirp = None

    def chirp(self):
        if self._chirp is None:
            from ria_b200 import sync
            self._chirp = sync.chirp_generate(device=self.device, ctx=self.ctx)
        return self._chirp

    def _rows(self, bodies):
        import torch
        pre = self.chirp()
        n, flen = bodies.shape
        rows = torch.zeros((n, LEAD + pre.numel() + flen + TAIL), dtype=torch.float32, device=self.device)
        rows[:, LEAD:LEAD + pre.numel()] = pre
        rows[:, LEAD + pre.numel():LEAD + pre.numel() + flen] = bodies
        return rows

    def get(self, rec):
        import torch
        from ria_b200 import fec, mcdpsk, ofdm, selection, sync, txsynth
        key = (rec.waveform, rec.modulation, rec.rate, rec.spreading)
        if key in self.modes:
            return self.modes[key]
        rng = np.random.default_rng(1000 + hash(key) % 1000)
        if rec.waveform == selection.MC_DPSK:
            bits = 1 if rec.modulation == 0 else 2
            cfg = mcdpsk.MultiCarrierDPSKConfig.default(bits, max(1, rec.spreading), rec.num_carriers or 10)
            k = fec.code_params(rec.rate)[0]
            info = rng.integers(0, 2, size=(POOL, k), dtype=np.uint8)
            cw = np.packbits(txsynth.ldpc_encode_bits(info, rec.rate), axis=1)
            bodies = mcdpsk.mcdpsk_tx_frames(cfg, torch.from_numpy(cw).to(self.device), self.ctx)
            chain = mcdpsk.McdpskRxChain(cfg, rec.rate, fec.recommended_iterations(rec.rate), 0.9375, 0.15, self.ctx)
            mode = dict(kind="mcdpsk", name=f"MC-DPSK {'DBPSK' if bits == 1 else 'DQPSK'} x{max(1, rec.spreading)} R{rec.rate}",
                        pool=self._rows(bodies), frame_len=bodies.shape[1], chain=chain, k=k,
                        dem=mcdpsk.MCDPSKDemodulator(cfg, self.ctx),
                        sent=torch.from_numpy(np.packbits(info, axis=1)).to(self.device))
        else:
            cfg = ofdm.ModemConfig.for_waveform(rec.modulation, rec.rate)
            bpc = fec.code_params(rec.rate)[0] // 8
            frames = txsynth.make_data_frames("K1ABC", "W2XYZ", 0, rng.integers(0, 256, size=(POOL, 4 * bpc - 19 - 2), dtype=np.uint8), bpc)
            fr_dev = torch.from_numpy(frames).to(self.device)
            coded = ofdm.encode_fixed_frame_batch(fr_dev, rec.rate, True, cfg.bitsPerSymbol(), self.ctx)
            bodies = ofdm.ofdm_tx_frames(cfg, coded, self.ctx)
            mode = dict(kind="ofdm", name=f"OFDM mod{rec.modulation} R{rec.rate}", pool=self._rows(bodies),
                        frame_len=bodies.shape[1], chain=ofdm.OfdmRxChain(cfg, rec.rate, True, self.ctx),
                        dem=ofdm.OFDMDemodulator(cfg, self.ctx), sent=fr_dev, sync=sync.ChirpSync(None, self.ctx))
        mode["body_start"] = LEAD + self.chirp().numel()
        self.modes[key] = mode
        return mode


def _channel(mode, snr_db, n, first_id, channel_cfg, seed, ctx, device):
    import torch
    from ria_b200 import sim
    snr = torch.full((n,), float(snr_db), dtype=torch.float32, device=device)
    return sim.watterson_batch(channel_cfg, mode["pool"], n, snr, seed=seed, first_frame_id=first_id, ctx=ctx)


def measure_fading(mode, snr_db, channel_cfg, seed, ctx, device, n=1024):
    """mean of the demodulator's own fading index (OFDMDemodulator / MultiCarrierDPSKDemodulator::getFadingIndex) over a
    probe batch through this channel: what the adaptive ladder is fed with (waveform_selection.hpp:112-222)"""
    rx = _channel(mode, snr_db, n, 1 << 40, channel_cfg, seed ^ 0x5bd1, ctx, device)
    b0 = mode["body_start"]
    out = (mode["dem"].process_presynced_batch if mode["kind"] == "ofdm" else mode["dem"].process_batch)(
        rx[:, b0:b0 + mode["frame_len"]].contiguous())
    return float(out["fading"].float().mean().item())


def run_point(mode, snr_db, n_local, first_id, channel_cfg, seed, ctx, device):
    """-> int64 tensor [frames, frames_ok, crc_ok_but_wrong, sync_miss] for this rank's shard.  Every frame goes through
    acquisition first (dual-chirp sync on the whole reception, as IWaveform::detectSync), is demodulated at the detected
    training start with the detected CFO and decoded; it counts as decoded only when its bytes are the bytes that were sent
    (the false-positive repair of decodeFixedFrame tries tens of thousands of bit flips against a 16-bit CRC, so
    CRC-valid frames with a wrong payload exist and are reported separately)."""
    import torch
    from ria_b200 import ofdm as ofdm_mod
    from ria_b200 import sync as rsync
    c = torch.zeros(4, dtype=torch.int64, device=device)
    flen = mode["frame_len"]
    for off in range(0, n_local, CHUNK):
        n = min(CHUNK, n_local - off)
        gid = first_id + off
        rx = _channel(mode, snr_db, n, gid, channel_cfg, seed, ctx, device)
        want = mode["sent"][(torch.arange(n, device=device) + gid) % POOL]
        window = min(120000, rx.shape[1])
        if mode["kind"] == "mcdpsk":
            out = mode["chain"].process_batch(rx, flen, window)
            det = out["sync"].view(torch.int32)[:, 0] != 0
            nbytes = (mode["k"] + 7) // 8
            same = (out["info"][:, :nbytes] == want[:, :nbytes]).all(dim=1)
            ok = out["ok"].bool() & det
        else:
            sy = mode["sync"].detect_dual_batch(rx[:, :window].contiguous(), 0.15)
            f = sy.view(torch.int32)
            det = f[:, 0] != 0
            start = torch.where(det, f[:, 7] + 28800, torch.zeros_like(f[:, 7])).clamp(0, rx.shape[1] - flen).long()
            cfo = torch.where(det, sy.view(torch.float32)[:, 3], torch.zeros(n, device=device))
            # OFDMChirpWaveform::process: CFO phase accumulated from the start of the audio (ofdm_chirp_waveform.cpp:404-413)
            ph = torch.remainder(-2.0 * np.pi * cfo.double() * start.double() / 48000.0 + np.pi, 2.0 * np.pi) - np.pi
            frames = torch.gather(rx, 1, start[:, None] + torch.arange(flen, device=device)[None, :])
            data, status, _ = mode["chain"].process_batch(frames, cfo.contiguous(), ph.float().contiguous())
            st = status.view(torch.uint8)
            o = {k: ofdm_mod.FRAME_STATUS_DTYPE.fields[k][1] for k in ("all_ok", "header_valid", "frame_crc_ok")}
            ok = (st[:, o["all_ok"]] == 1) & (st[:, o["header_valid"]] == 1) & (st[:, o["frame_crc_ok"]] == 1) & det
            same = (data[:, : want.shape[1]] == want).all(dim=1)
        c += torch.stack([torch.tensor(n, device=device), (ok & same).sum(), (ok & ~same).sum(), (~det).sum()]).to(torch.int64)
    return c


def run_sweep(args):
    import torch
    import ria_b200
    from ria_b200 import dist as rdist
    from ria_b200 import selection, sim
    rank, world, local = rdist.env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        rdist.init("nccl", device)
    ctx = ria_b200.Context(local)
    if not getattr(args, "first_pass_only", False):
        ctx.set_decode_flags(ria_b200.DECODE_FULL)      # the reference's complete decodeFixedFrame (ladder + repair)
    lo, n_local = rdist.shard_range(args.frames, rank, world)
    cache = ModeCache(ctx, device)
    results = []
    fading = FADING_INDEX[args.condition]               # first guess; replaced by the demodulator's estimate point by point
    measure = not getattr(args, "fixed_fading", False)
    t_all = time.perf_counter()
    for snr_db in np.arange(args.snr_min, args.snr_max + 1e-9, args.snr_step):
        channel_cfg = sim.WattersonConfig.preset(CONDITIONS[args.condition], float(snr_db))
        channel_cfg.stationary_start = 1
        seed = 7000 + int(round(snr_db * 10))
        rec = selection.recommendWaveformAndRate(float(snr_db), fading)
        mode = cache.get(rec)
        if measure:
            # the ladder is fed with what the demodulator of the current mode reports on this channel; one re-selection
            fading = measure_fading(mode, snr_db, channel_cfg, seed, ctx, device)
            if world > 1:
                fading = rdist.max_over_ranks(fading, device)       # every rank must take the same branch
            rec = selection.recommendWaveformAndRate(float(snr_db), fading)
            mode = cache.get(rec)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = run_point(mode, snr_db, n_local, lo, channel_cfg, seed, ctx, device)
        if world > 1:
            c = rdist.allreduce_counters(c)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        frames, ok, wrong, miss = int(c[0]), int(c[1]), int(c[2]), int(c[3])
        line = {"snr_db": float(snr_db), "mode": mode["name"], "fading_index_fed_to_ladder": round(fading, 4), "frames": frames,
                "frames_ok": ok, "crc_ok_but_wrong_payload": wrong, "sync_miss": miss, "fer": 1.0 - ok / max(1, frames),
                "estimated_throughput_bps": float(rec.estimated_throughput_bps), "frames_per_s": frames / dt}
        results.append(line)
        if rank == 0 and not args.quiet:
            print(json.dumps(line), flush=True)
    if rank == 0 and not args.quiet:
        print(json.dumps({"summary": "adaptive waveform sweep", "condition": args.condition, "n_gpus": world,
                          "frames_per_point": args.frames, "points": len(results),
                          "chain": "Watterson channel -> dual-chirp acquisition -> demodulation at the detected start / CFO -> "
                                   "decode (complete decodeFixedFrame unless --first-pass-only) -> payload check",
                          "seconds": time.perf_counter() - t_all}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    ctx.close()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4096, help="frames per SNR point (whole job)")
    ap.add_argument("--condition", default="moderate", choices=sorted(CONDITIONS))
    ap.add_argument("--snr-min", type=float, default=-14.0)
    ap.add_argument("--snr-max", type=float, default=30.0)
    ap.add_argument("--snr-step", type=float, default=2.0)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--fixed-fading", action="store_true",
                    help="feed the ladder the preset fading index of the condition instead of the demodulator's estimate")
    ap.add_argument("--first-pass-only", action="store_true",
                    help="skip decodeFixedFrame's retry ladder and false-positive repair")
    run_sweep(ap.parse_args())


if __name__ == "__main__":
    main()
