"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

Run here (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The reference ships no known-answer vectors of its own, so these are outputs
of the reference itself on seeded inputs; the GPU box checks against them without the
reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.bindings import (Ref, R1_4, R1_2, R2_3, R3_4, R5_6, RATE_K, RATE_MAX_ITER,  # noqa: E402
                             awgn_llrs, unpack_bits)


def ldpc_golden(ref, out_dir):
    rng = np.random.default_rng(20261018)
    # Es/N0 per rate chosen so that the set mixes early converging, late converging and failing cws
    esn0 = {R1_4: -2.5, R1_2: 1.8, R2_3: 4.3, R3_4: 5.3, R5_6: 6.5}
    n_cw = 48
    out = {}
    for rate in (R1_4, R1_2, R2_3, R3_4, R5_6):
        k = RATE_K[rate]
        llr = np.zeros((n_cw, 648), np.float32)
        data = rng.integers(0, 256, size=(n_cw, k // 8), dtype=np.uint8)
        for i in range(n_cw):
            cw = ref.ldpc_encode(rate, data[i])[:81]
            llr[i] = awgn_llrs(unpack_bits(cw), esn0[rate] + (i % 3) * 0.7, rng)
        # a few adversarial rows: zeros, huge values, -0.0, inf, NaN
        llr[40] = 0.0
        llr[41] = np.where(rng.random(648) < 0.5, 1e6, -1e6).astype(np.float32)
        llr[42, ::7] = -0.0
        llr[43, 5] = np.inf
        llr[43, 9] = -np.inf
        llr[44, 3] = np.nan
        llr[45] = np.abs(llr[45])          # all-zero codeword
        out[f"r{rate}_llr"] = llr
        out[f"r{rate}_data"] = data
        for tag, factor in (("a", 0.75), ("b", 0.9375)):
            info, ok, iters = ref.ldpc_decode_batch(rate, llr, RATE_MAX_ITER[rate], factor, 68)
            out[f"r{rate}_{tag}_info"] = info
            out[f"r{rate}_{tag}_ok"] = ok
            out[f"r{rate}_{tag}_iters"] = iters
        # H structure as seen through the encoder: column j of H_data = parity of unit vector e_j
        cols = np.zeros((k, 648 - k), np.uint8)
        for j in range(k):
            d = np.zeros((k + 7) // 8, np.uint8)
            d[j // 8] = 0x80 >> (j % 8)
            cols[j] = unpack_bits(ref.ldpc_encode(rate, d)[:81])[k:]
        out[f"r{rate}_hdata_cols"] = np.packbits(cols, axis=1)
    np.savez_compressed(os.path.join(out_dir, "ldpc_golden.npz"), **out)
    print("ldpc_golden.npz written:", {k: v.shape for k, v in out.items() if k.endswith("iters")})


def ofdm_golden(ref, out_dir):
    from tests.ofdm_common import CASES, apply_cfo, awgn, make_cfg, tx_frame
    out = {}
    for name, mod, spacing, use_pilots, rate, snr_db in CASES:
        cfg = make_cfg(mod, spacing, use_pilots)
        rng = np.random.default_rng(abs(hash(name)) % (1 << 31) if False else sum(map(ord, name)))
        rxs, softs, snrs, cfo_out, cfos, phases = [], [], [], [], [], []
        for i in range(3):
            tx, _, _ = tx_frame(ref, cfg, rate, rng, seq=i)
            cfo = (0.0, 1.7, -3.2)[i]
            rx = awgn(apply_cfo(tx, cfo) if cfo else tx, snr_db, rng)
            est = np.float32(cfo + (0.0, 0.1, 1.2)[i])
            ph = np.float32((0.0, 0.5, -2.0)[i])
            r = ref.ofdm_process_presynced(cfg, rx, float(est), float(ph))
            rxs.append(rx.astype(np.float16).astype(np.float32))   # store quantised samples (smaller file)
            r = ref.ofdm_process_presynced(cfg, rxs[-1], float(est), float(ph))
            softs.append(r["soft"]); snrs.append(r["snr_db"]); cfo_out.append(r["cfo"])
            cfos.append(est); phases.append(ph)
        out[f"{name}_rx"] = np.stack(rxs).astype(np.float16)
        out[f"{name}_soft"] = np.stack(softs)
        out[f"{name}_snr"] = np.array(snrs, np.float32)
        out[f"{name}_cfo_out"] = np.array(cfo_out, np.float32)
        out[f"{name}_cfo"] = np.array(cfos, np.float32)
        out[f"{name}_phase"] = np.array(phases, np.float32)
    np.savez_compressed(os.path.join(out_dir, "ofdm_golden.npz"), **out)
    print("ofdm_golden.npz written")


def frame_golden(ref, out_dir):
    """Complete v2::decodeFixedFrame (first pass + retry ladder + false-positive repair) on degraded
    frames, plus the burst de-interleaver: soft bits in, the reference's decoded flags / bytes out."""
    from oracle.bindings import BYTES_PER_CW
    rng = np.random.default_rng(20261019)
    out = {}
    for name, rate, esn0, bps in (("r14", R1_4, -1.0, 53), ("r12", R1_2, 2.6, 106), ("r23", R2_3, 4.6, 176), ("r34", R3_4, 7.0, 264)):
        bpc = BYTES_PER_CW[rate]
        n = 20
        soft = np.empty((n, 2592), np.float16)
        data = np.empty((n, 4 * bpc), np.uint8)
        ok = np.empty((n, 4), np.uint8)
        for i in range(n):
            payload = rng.integers(0, 256, size=4 * bpc - 19 - int(rng.integers(0, 8)), dtype=np.uint8)
            if i == 3:
                payload[bpc - 17] = 0xD5                     # DATA_CW_MARKER quirk of reassemble() (frame_v2.cpp:974)
            frame = ref.make_data_frame("K1ABC", "W2XYZ", i, payload)
            coded = ref.encode_fixed_frame(frame, rate, True, bps)
            llr = awgn_llrs(np.unpackbits(coded)[:2592], esn0 + (i % 4) * 0.5, rng)
            soft[i] = llr.astype(np.float16)                 # stored as fp16 (exactly representable inputs keep the file small)
            data[i], ok[i] = ref.decode_fixed_frame_full(soft[i].astype(np.float32), rate, True, bps)
        out[f"{name}_soft"], out[f"{name}_data"], out[f"{name}_ok"] = soft, data, ok
        out[f"{name}_rate"], out[f"{name}_bps"] = np.int32(rate), np.int32(bps)
        print(name, "frames decoded by the reference:", int(ok.all(axis=1).sum()), "of", n)
    phys = rng.standard_normal((4, 2592)).astype(np.float16)
    out["burst_physical"] = phys
    out["burst_logical"] = ref.burst_deinterleave(phys.astype(np.float32)).astype(np.float16)
    np.savez_compressed(os.path.join(out_dir, "frame_golden.npz"), **out)
    print("frame_golden.npz written")


def cox_golden(ref, out_dir):
    """OFDM_COX acquisition (OFDMDemodulator::searchForSync): windows in, found / LTS position / CFO / noise floor out,
    for two thresholds and a carried noise floor, plus Impl::measureCorrelation at a few offsets."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_ofdm_cox_gpu import _windows            # the same window generator the live-reference test uses
    from tests.ofdm_common import make_cfg, QAM64, DQPSK
    out = {}
    for name, mod, spacing, pilots, rate in (("qam64_sp4", QAM64, 4, 1, R3_4), ("dqpsk_sp5", DQPSK, 5, 1, R1_2)):
        cfg = make_cfg(mod, spacing, pilots)
        rng = np.random.default_rng(sum(map(ord, name)))
        window = 20000
        wins, _ = _windows(ref, rng, 8, window, cfg, rate)
        x = np.stack(wins).astype(np.float16)              # stored quantised: the expected values are computed on these
        xf = x.astype(np.float32)
        out[f"{name}_win"] = x
        for tag, thr, nf_in in (("a", 0.8, 0.0), ("b", 0.6, 0.002)):
            res = [ref.ofdm_cox_search_sync(cfg, xf[i], thr, nf_in) for i in range(len(xf))]
            out[f"{name}_{tag}_thr"] = np.float32(thr)
            out[f"{name}_{tag}_nf_in"] = np.float32(nf_in)
            out[f"{name}_{tag}_found"] = np.array([r[0] for r in res], np.uint8)
            out[f"{name}_{tag}_pos"] = np.array([r[1] if r[0] else -1 for r in res], np.int64)
            out[f"{name}_{tag}_cfo"] = np.array([r[2] if r[0] else 0.0 for r in res], np.float32)
            out[f"{name}_{tag}_nf_out"] = np.array([r[3] for r in res], np.float32)
        offs = np.array([0, 64, 1000, 2048, 4096, 6000, 7777, 9000], np.int32)
        out[f"{name}_corr_off"] = offs
        out[f"{name}_corr"] = np.array([ref.ofdm_cox_correlation(cfg, xf[i], int(offs[i])) for i in range(len(xf))], np.float32)
        print(name, "found:", out[f"{name}_a_found"].tolist(), out[f"{name}_b_found"].tolist())
    np.savez_compressed(os.path.join(out_dir, "cox_golden.npz"), **out)
    print("cox_golden.npz written")


if __name__ == "__main__":
    ref = Ref()
    if "cox" in sys.argv[1:] or len(sys.argv) == 1:
        cox_golden(ref, HERE)
    if "frame" in sys.argv[1:] or len(sys.argv) == 1:
        frame_golden(ref, HERE)
    if "ldpc" in sys.argv[1:] or len(sys.argv) == 1:
        ldpc_golden(ref, HERE)
    if "ofdm" in sys.argv[1:] or len(sys.argv) == 1:
        ofdm_golden(ref, HERE)
