#!/usr/bin/env python
"""Adaptive-waveform SNR sweep over a Watterson channel (BASELINE.json configs[4]).

For every SNR point the reference's selection ladder (protocol::recommendWaveformAndRate,
src/protocol/waveform_selection.hpp:112-222) picks waveform, modulation, code rate and spreading;
the transmissions are synthesised once per mode on the host, the HF channel
(sim::WattersonChannel, src/sim/hf_channel.hpp) runs on the device for every frame, the frames go
through the batched receive chain of that mode, and the frame-error counters are summed over the
GPUs with one NCCL all-reduce per SNR point.  Frames are sharded over ranks; global frame ids key
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


class ModeCache:
    """TX pools and receive chains per (waveform, modulation, rate, spreading)."""

    def __init__(self, ctx, device):
        self.ctx, self.device, self.modes = ctx, device, {}

    def get(self, rec):
        import torch
        from ria_b200 import fec, mcdpsk, ofdm, selection, txsynth
        key = (rec.waveform, rec.modulation, rec.rate, rec.spreading)
        if key in self.modes:
            return self.modes[key]
        rng = np.random.default_rng(1000 + hash(key) % 1000)
        if rec.waveform == selection.MC_DPSK:
            bits = 1 if rec.modulation == 0 else 2
            cfg = mcdpsk.MultiCarrierDPSKConfig.default(bits, max(1, rec.spreading), rec.num_carriers or 10)
            k = fec.code_params(rec.rate)[0]
            rows, sent = [], []
            for _ in range(POOL):
                info = rng.integers(0, 2, size=k, dtype=np.uint8)
                cw = np.packbits(txsynth.ldpc_encode_bits(info, rec.rate))
                rows.append(txsynth.mcdpsk_modulate_frame(cfg, cw.tobytes()))
                sent.append(np.packbits(info))
            dem = mcdpsk.MCDPSKDemodulator(cfg, self.ctx)
            dec = fec.LDPCDecoder(rec.rate, self.ctx)
            dec.setMaxIterations(fec.recommended_iterations(rec.rate))
            dec.setMinSumFactor(0.9375)
            mode = dict(kind="mcdpsk", name=f"MC-DPSK {'DBPSK' if bits == 1 else 'DQPSK'} x{max(1, rec.spreading)} R{rec.rate}",
                        pool=torch.from_numpy(np.stack(rows)).to(self.device),
                        sent=torch.from_numpy(np.stack(sent)).to(self.device), dem=dem, dec=dec, k=k)
        else:
            cfg = ofdm.ModemConfig.for_waveform(rec.modulation, rec.rate)
            pool, raw = txsynth.make_frame_pool(cfg, rec.rate, POOL, seed=int(rng.integers(1 << 30)))
            chain = ofdm.OfdmRxChain(cfg, rec.rate, True, self.ctx)
            sent = np.stack([np.frombuffer(fr, dtype=np.uint8) for fr in raw])        # every frame is 4 x bytes_per_cw long
            mode = dict(kind="ofdm", name=f"OFDM mod{rec.modulation} R{rec.rate}",
                        pool=torch.from_numpy(pool).to(self.device), chain=chain,
                        sent=torch.from_numpy(sent).to(self.device))
        self.modes[key] = mode
        return mode


def run_point(mode, snr_db, n_local, first_id, channel_cfg, seed, ctx, device):
    """-> int64 tensor [frames, frames_ok, crc_ok_but_wrong] for this rank's shard.  A frame counts as decoded
    only when its bytes are the bytes that were sent: the false-positive repair of decodeFixedFrame tries tens
    of thousands of bit flips against a 16-bit CRC, so CRC-valid frames with a wrong payload exist and are
    reported separately."""
    import torch
    from ria_b200 import sim
    snr = torch.full((n_local,), float(snr_db), dtype=torch.float32, device=device)
    rx = sim.watterson_batch(channel_cfg, mode["pool"], n_local, snr, seed=seed, first_frame_id=first_id, ctx=ctx)
    if mode["kind"] == "ofdm":
        from ria_b200 import ofdm as ofdm_mod
        data, status, _ = mode["chain"].process_batch(rx)
        st = status.view(torch.uint8)
        off = {k: ofdm_mod.FRAME_STATUS_DTYPE.fields[k][1] for k in ("all_ok", "header_valid", "frame_crc_ok")}
        crc_ok = (st[:, off["all_ok"]] == 1) & (st[:, off["header_valid"]] == 1) & (st[:, off["frame_crc_ok"]] == 1)
        want = mode["sent"][(torch.arange(n_local, device=device) + first_id) % POOL]
        same = (data[:, : want.shape[1]] == want).all(dim=1)
        return torch.stack([torch.tensor(n_local, device=device), (crc_ok & same).sum(), (crc_ok & ~same).sum()]).to(torch.int64)
    out = mode["dem"].process_batch(rx)
    info, ok, _ = mode["dec"].decode_batch(out["llr"][:, :648].contiguous())
    nbytes = (mode["k"] + 7) // 8
    want = mode["sent"][(torch.arange(n_local, device=device) + first_id) % POOL]
    same = (info[:, :nbytes] == want[:, :nbytes]).all(dim=1)
    return torch.stack([torch.tensor(n_local, device=device), (ok.bool() & same).sum(), (ok.bool() & ~same).sum()]).to(torch.int64)


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
    t_all = time.perf_counter()
    for snr_db in np.arange(args.snr_min, args.snr_max + 1e-9, args.snr_step):
        rec = selection.recommendWaveformAndRate(float(snr_db), FADING_INDEX[args.condition])
        mode = cache.get(rec)
        channel_cfg = sim.WattersonConfig.preset(CONDITIONS[args.condition], float(snr_db))
        channel_cfg.stationary_start = 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = run_point(mode, snr_db, n_local, lo, channel_cfg, 7000 + int(round(snr_db * 10)), ctx, device)
        if world > 1:
            c = rdist.allreduce_counters(c)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        frames, ok, wrong = int(c[0]), int(c[1]), int(c[2])
        line = {"snr_db": float(snr_db), "mode": mode["name"], "frames": frames, "frames_ok": ok,
                "crc_ok_but_wrong_payload": wrong, "fer": 1.0 - ok / max(1, frames), "estimated_throughput_bps": float(rec.estimated_throughput_bps),
                "frames_per_s": frames / dt}
        results.append(line)
        if rank == 0 and not args.quiet:
            print(json.dumps(line), flush=True)
    if rank == 0 and not args.quiet:
        print(json.dumps({"summary": "adaptive waveform sweep", "condition": args.condition, "n_gpus": world,
                          "frames_per_point": args.frames, "points": len(results),
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
    ap.add_argument("--first-pass-only", action="store_true",
                    help="skip decodeFixedFrame's retry ladder and false-positive repair")
    run_sweep(ap.parse_args())


if __name__ == "__main__":
    main()
