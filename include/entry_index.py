#!/usr/bin/env python
"""Prints the entry-point index of INTEGRATION.md from the comments of ria_b200.h:
every exported function, the first sentence of its comment and the reference locations it cites.

    python include/entry_index.py > /tmp/index.md
"""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(HERE, "ria_b200.h")).read()

# a comment block immediately followed by one or more declarations (the `_host` twins share the comment of their `_dev`)
pat = re.compile(r"/\*(?P<c>(?:[^*]|\*(?!/))*)\*/\s*(?P<d>(?:(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\*?\s+ria_[a-z0-9_]+\s*\([^;]*\);\s*)+)")
cite = re.compile(r"[A-Za-z0-9_/]+\.(?:cpp|hpp|h):\d+(?:-\d+)?|(?<![A-Za-z0-9_.]):\d+(?:-\d+)?")
rows = []
for m in pat.finditer(src):
    comment = " ".join(line.strip(" *") for line in m.group("c").strip().splitlines())
    comment = re.sub(r"\s+", " ", comment).strip()
    if comment.startswith("----"):                       # a section banner, not a description
        comment = comment.strip("- ").strip() + " (section banner; see the declarations)"
    first = re.split(r"(?<=[a-z0-9\)])[.:;] (?=[A-Z`])", comment, maxsplit=1)[0]
    if len(first) > 170:
        first = first[:167].rsplit(" ", 1)[0] + " ..."
    cites = []
    for c in cite.findall(comment):
        if c not in cites:
            cites.append(c)
    names = re.findall(r"(ria_[a-z0-9_]+)\s*\(", m.group("d"))
    for n in names:
        rows.append((n, first, ", ".join(cites[:4])))

# declarations that share a banner or sit behind a typedef: stated by hand
OVERRIDE = {
    "ria_ctx_create": ("One context per GPU (device memory, streams, tables); fails when there is no usable sm_100 GPU: no CPU fallback", "one waveform / decoder object per thread, src/gui/modem/streaming_decoder.cpp:718-723"),
    "ria_ctx_destroy": ("Releases everything the context owns", ""),
    "ria_ctx_set_stream": ("Bind the context to an existing cudaStream_t (e.g. torch's current stream)", ""),
    "ria_ctx_synchronize": ("Wait for the context's stream", ""),
    "ria_last_error": ("Text of the last error on this context", ""),
    "ria_version": ("Library version string", ""),
    "ria_ctx_set_timing": ("Per-kernel timing for bench.py (CUDA events around every launch)", ""),
    "ria_ctx_get_timing": ("Summed elapsed time and launch count of one kernel kind since timing was enabled", ""),
    "ria_ctx_set_decode_flags": ("RIA_DECODE_RETRY_LADDER / RIA_DECODE_FP_REPAIR / RIA_DECODE_FULL: which phases of v2::decodeFixedFrame the frame entry points run", "src/protocol/frame_v2.cpp:1335-1385, :1389-1546, :1558-1916"),
    "ria_ctx_get_decode_flags": ("Current decode flags", ""),
    "ria_chirp_config_default": ("ChirpConfig as both waveforms set it up (300-2700 Hz, 500 ms, 100 ms gap)", "src/sync/chirp_sync.hpp:29-38"),
    "ria_ofdm_symbol_samples": ("Samples per OFDM symbol of a config (ModemConfig::getSymbolDuration)", ""),
    "ria_ofdm_data_carriers": ("Data carriers of a config", ""),
    "ria_ofdm_pilot_carriers": ("Pilot carriers of a config", ""),
    "ria_chirp_detect_dual_batch_host": ("HOST-buffer twin of ria_chirp_detect_dual_batch_dev: IWaveform::detectSync of both chirp waveforms", "ofdm_chirp_waveform.cpp:163-205, mc_dpsk_waveform.cpp:177-224"),
    "ria_zc_detect_batch_host": ("HOST-buffer twin of ria_zc_detect_batch_dev: MCDPSKWaveform::detectDataSync", "mc_dpsk_waveform.cpp:227-292"),
    "ria_ofdm_data_sync_batch_host": ("HOST-buffer twin of ria_ofdm_data_sync_batch_dev: OFDMChirpWaveform::detectDataSync", "ofdm_chirp_waveform.cpp:207-384"),
    "ria_mcdpsk_process_batch_host": ("HOST-buffer twin of ria_mcdpsk_process_batch_dev: MCDPSKWaveform::process", "mc_dpsk_waveform.cpp:294-338"),
    "ria_ofdm_cox_search_sync_batch_host": ("HOST-buffer twin of ria_ofdm_cox_search_sync_batch_dev: OFDMNvisWaveform::detectSync (noise floor a host array too)", "ofdm_cox_waveform.cpp:121-153"),
    "ria_nccl_get_unique_id": ("ncclGetUniqueId for the helper below (rank 0)", ""),
    "ria_nccl_comm_create": ("ncclCommInitRank on the context's device (NCCL resolved from the process image)", ""),
    "ria_nccl_comm_destroy": ("ncclCommDestroy", ""),
    "ria_counters_allreduce": ("The path's only collective: ncclAllReduce(sum, int64) of a counter vector, in place, on the context stream", "tools/cli_simulator.cpp:2226-2290 (the per-station statistics it sums)"),
}
have = {n for n, _, _ in rows}
for n in OVERRIDE:
    if n not in have:
        rows.append((n, "", ""))
rows = [(n,) + (OVERRIDE[n] if n in OVERRIDE else (f, c)) for n, f, c in rows]
order = re.findall(r"\b(ria_[a-z0-9_]+)\s*\(", src)
rows.sort(key=lambda r: order.index(r[0]))

print("| entry point | what it is (first sentence of its comment in `include/ria_b200.h`) | reference locations cited there |")
print("|---|---|---|")
for n, first, cites in rows:
    print(f"| `{n}` | {first.replace('|', '/')} | {cites or '—'} |")
