#!/usr/bin/env python
"""bench.py -- the measurement contract (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one pass of the receive hot path over one batch of synthetic input that is already
resident in HBM.  Rank 0 prints ONE JSON line.  For N > 1 the driver launches this file under
torch.distributed.run (one rank per GPU, NCCL); frames are sharded across ranks (weak scaling:
fixed per-GPU batch) and the only collective is an all-reduce of the error counters.
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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------
# workload: batched LDPC decode (BASELINE.json configs[1])
# ---------------------------------------------------------------------------------------------

class LdpcWorkload:
    """configs[1]: R1/4..R3/4 min-sum decode of 1M codewords per rate from synthetic AWGN LLRs
    (LLR model of tools/test_chase_cache.cpp:20-34; Es/N0 and decoder settings of SURVEY.md 8d)."""

    name = "ldpc_r14_r34_1M_cw_per_rate"
    RATES = (0, 2, 3, 4)                       # R1/4, R1/2, R2/3, R3/4
    ESN0 = {0: 1.0, 2: 4.0, 3: 6.0, 4: 7.0}
    MAX_ITER = {0: 50, 2: 80, 3: 70, 4: 60}
    FACTOR = 0.9375                            # decodeFixedFrame / robustDecodeSingleCW setting
    CW_PER_FRAME = 4                           # v2 fixed frame = 4 codewords

    def __init__(self, n_cw: int):
        self.n_cw = n_cw

    # -- synthetic input: encode with the library's own H (systematic: parity = H_data . info) --
    def _codeword_bits(self, rate, n_distinct, rng):
        from ria_b200 import fec
        k, m, _ = fec.code_params(rate)
        row_ptr, edge_var = fec.get_matrix(rate)
        info = rng.integers(0, 2, size=(n_distinct, k), dtype=np.uint8)
        par = np.zeros((n_distinct, m), np.uint8)
        for i in range(m):
            vs = edge_var[row_ptr[i]:row_ptr[i + 1] - 1]
            par[:, i] = info[:, vs].sum(axis=1) & 1
        return np.concatenate([info, par], axis=1)

    def make_llr_host(self, rate, n, seed):
        rng = np.random.default_rng(seed)
        bits = self._codeword_bits(rate, 64, rng)
        snr = np.float32(10 ** (self.ESN0[rate] / 10))
        s = 1.0 - 2.0 * bits[rng.integers(0, 64, size=n)].astype(np.float32)
        noise = rng.standard_normal(s.shape, dtype=np.float32) / np.sqrt(snr)
        return (2.0 * (s + noise) * snr).astype(np.float32), bits

    def setup(self, ctx, device, rank):
        import torch
        from ria_b200 import fec
        self.torch = torch
        self.ctx = ctx
        self.dec = {}
        self.llr = {}
        self.info_ref = {}
        self.prot_mask = {}
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
        out = {}
        for rate in self.RATES:
            out[rate] = self.dec[rate].decode_batch(self.llr[rate])
        self.out = out

    def launches_per_step(self):
        return len(self.RATES)

    def units_per_step(self):      # frames (4 codewords each)
        return len(self.RATES) * self.n_cw / self.CW_PER_FRAME

    def samples_per_step(self):
        return 0.0

    def algorithmic_bytes_per_step(self):
        from ria_b200 import fec
        b = 0
        for rate in self.RATES:
            k = fec.code_params(rate)[0]
            b += self.n_cw * (648 * 4 + (k + 7) // 8 + 1 + 4)
        return b

    def counters(self):
        """[cw, cw_fail, info_byte_errors_among_ok, sum_iters]"""
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

    # -- end to end: host buffers through the *_host C ABI entry --
    def setup_e2e(self, n_e2e):
        self.e2e_n = n_e2e
        self.e2e_llr = {}
        self._pinned = []
        for rate in self.RATES:
            host, _ = self.make_llr_host(rate, n_e2e, 555 + rate)
            pin = self.torch.empty(host.shape, dtype=self.torch.float32, pin_memory=True)
            pin.numpy()[:] = host
            self._pinned.append(pin)
            self.e2e_llr[rate] = pin.numpy()

    def step_e2e(self):
        res = {}
        for rate in self.RATES:
            res[rate] = self.dec[rate].decode_batch_host(self.e2e_llr[rate])
        return res

    def e2e_units(self):
        return len(self.RATES) * self.e2e_n / self.CW_PER_FRAME

    def e2e_bytes(self):
        from ria_b200 import fec
        h2d = len(self.RATES) * self.e2e_n * 648 * 4
        d2h = sum(self.e2e_n * ((fec.code_params(r)[0] + 7) // 8 + 1 + 4) for r in self.RATES)
        return h2d, d2h

    # -- CPU baseline: the unmodified reference, one process per host core --
    def cpu_sample(self, per_rate):
        return {rate: self.make_llr_host(rate, per_rate, 777 + rate)[0] for rate in self.RATES}


def _cpu_ldpc_worker(args):
    """One process per core: synthesise its own sample (untimed), then time the decode."""
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


def cpu_baseline_ldpc(wl: LdpcWorkload, per_rate_per_core: int):
    """Times the reference's own LDPCDecoder on the host cores: every core decodes
    per_rate_per_core codewords of each rate (one process per core; the reference is not
    thread-safe).  Throughput = sum over cores of (codewords / busy time)."""
    import multiprocessing as mp
    from oracle.bindings import Ref
    kind = "reference" if Ref.available() else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    jobs = [(per_rate_per_core, 1000 + 17 * c, kind) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        busy = pool.map(_cpu_ldpc_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    per_core_cw = len(wl.RATES) * per_rate_per_core
    rate_sum = sum(per_core_cw / b for b in busy)
    return {"value": rate_sum / wl.CW_PER_FRAME, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{per_rate_per_core} codewords per rate per core x {cores} cores "
                      f"({cores * per_core_cw} codewords), {wall:.1f} s wall, "
                      f"{max(busy):.1f} s max busy per core"}


# ---------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------

def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    wl = LdpcWorkload(args.batch)
    per = max(64, args.cpu_sample)
    vals = []
    for _ in range(args.warmup):
        cpu_baseline_ldpc(wl, max(16, per // 8))
    for _ in range(args.steps):
        vals.append(cpu_baseline_ldpc(wl, per))
    v = float(np.mean([x["value"] for x in vals]))
    last = vals[-1]
    line = {
        "impl": "reference", "metric": "decoded_frames_per_s", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (last["cores"] * len(wl.RATES) * per / wl.CW_PER_FRAME) / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl.name, "note": "reference LDPCDecoder::decodeSoft on host cores, "
                   "bounded sample of the same workload"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ria_b200", choices=["ria_b200", "reference"])
    ap.add_argument("--workload", default="ldpc")
    ap.add_argument("--batch", type=int, default=1 << 20, help="codewords per rate per GPU")
    ap.add_argument("--e2e-batch", type=int, default=1 << 17)
    ap.add_argument("--cpu-sample", type=int, default=4096, help="codewords per rate per core")
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
    wl = LdpcWorkload(args.batch)
    with torch.cuda.stream(stream):
        wl.setup(ctx, device, rank)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(max(args.warmup, 3)):
            wl.step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
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
        clocks = sampler.stop() if rank == 0 else None

        # error counters: the only cross-GPU exchange of the path (one all-reduce)
        cnt = wl.counters()
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        ms_per_step = ms_max / args.steps
        units = wl.units_per_step() * world
        value = units / (ms_per_step * 1e-3)

        # end to end through the host-buffer C ABI (H2D + kernels + D2H inside the timed region)
        wl.setup_e2e(args.e2e_batch)
        wl.step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            wl.step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = wl.e2e_units() * world / float(te.item())
        h2d, d2h = wl.e2e_bytes()

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg = wl.algorithmic_bytes_per_step() / wl.launches_per_step()
        kern_s = (ms / args.steps) * 1e-3 / wl.launches_per_step()
        achieved = alg / kern_s / 1e9
        c = cnt.cpu().numpy()
        line = {
            "metric": "decoded_frames_per_s", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "codewords_per_rate_per_gpu": args.batch,
                       "rates": "R1/4,R1/2,R2/3,R3/4", "esn0_db": "1,4,6,7",
                       "min_sum_factor": wl.FACTOR, "max_iter": "50,80,70,60",
                       "frame": "4 codewords (v2 fixed frame)",
                       "l2": "inputs (2.7 GB per rate) exceed the 126 MB L2; no flush needed"},
            "codewords_per_s": value * wl.CW_PER_FRAME,
            "counters": {"codewords": int(c[0]), "cw_fail": int(c[1]),
                         "ok_but_wrong_protected_bits": int(c[2]), "mean_iters": float(c[3]) / max(1, int(c[0]))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "ldpc_decode_kernel", "note": "LDPC is SM/shared-memory bound "
                         "by design (SURVEY 8d); HBM fraction reported for the contract"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "batch": args.e2e_batch},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_ldpc(wl, args.cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
