"""GPU: batched dual-chirp detection vs sync::ChirpSync::detectDualChirp of the unmodified
reference: raw and CFO-corrected peak positions identical, correlations within 1e-4."""
import numpy as np
import pytest

from tests.ofdm_common import apply_cfo

pytestmark = pytest.mark.gpu


def _windows(ref, n, window, rng, snrs, cfo_max=0.0, empty=0.15):
    chirp = ref.chirp_generate()                        # 57 600 samples: up, gap, down, gap
    wins = []
    for i in range(n):
        w = np.zeros(window, np.float32)
        if rng.random() >= empty:
            pos = int(rng.integers(0, max(1, window - len(chirp) - 2000)))
            sig = chirp
            if cfo_max:
                sig = apply_cfo(chirp, float(rng.uniform(-cfo_max, cfo_max)))
            w[pos:pos + len(sig)] += sig
            tail = window - pos - len(sig)
            w[pos + len(sig):] += 0.25 * np.sin(2 * np.pi * 900 * np.arange(tail) / 48000).astype(np.float32)
        snr = float(snrs[i % len(snrs)])
        w += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(0.125 / 10 ** (snr / 10)))
        wins.append(w)
    return wins


def _compare(ref, wins, got, thr):
    n_det = 0
    for i, w in enumerate(wins):
        r = ref.chirp_detect_dual(w, thr)
        g = got[i]
        assert g["detected"] == r.detected, (i, g, r.detected)
        assert abs(g["correlation"] - r.correlation) <= 1e-4 * max(1.0, r.correlation), i
        if r.detected:
            assert g["start_sample"] == r.start_sample and g["aux"] == r.aux, (i, g, r.start_sample, r.aux)
            assert abs(g["cfo_hz"] - r.cfo_hz) < 1e-4
            assert abs(g["snr_estimate"] - r.snr_estimate) <= 1e-4 * max(1.0, r.snr_estimate)
        n_det += r.detected
    return n_det


def test_matches_reference_production_window(ctx, ref):
    """StreamingDecoder's disconnected-mode search window: 120 000 samples
    (src/gui/modem/streaming_decoder.cpp:408-411)."""
    import torch
    from ria_b200 import sync
    rng = np.random.default_rng(11)
    wins = _windows(ref, 16, 120000, rng, [-12, -8, 0, 10, 20], cfo_max=40.0)
    cs = sync.ChirpSync(ctx=ctx)
    out = sync.results(cs.detect_dual_batch(torch.from_numpy(np.stack(wins)).cuda(), 0.15))
    assert _compare(ref, wins, out, 0.15) >= 8


def test_other_windows_and_edges(ctx, ref):
    import torch
    from ria_b200 import sync
    rng = np.random.default_rng(5)
    cs = sync.ChirpSync(ctx=ctx)
    wins = _windows(ref, 8, 70400, rng, [-5, 5, 15], cfo_max=10.0, empty=0.0)     # BASELINE.md probe size
    out = sync.results(cs.detect_dual_batch(torch.from_numpy(np.stack(wins)).cuda(), 0.15))
    _compare(ref, wins, out, 0.15)
    # too short for a dual chirp: default result
    short = sync.results(cs.detect_dual_batch(torch.zeros((2, 50000), device="cuda")))
    assert (short["detected"] == 0).all() and (short["start_sample"] == -1).all()
    # all-zero window: nothing found, correlation 0
    z = sync.results(cs.detect_dual_batch(torch.zeros((1, 120000), device="cuda")))
    r = ref.chirp_detect_dual(np.zeros(120000, np.float32))
    assert z["detected"][0] == r.detected == 0 and z["correlation"][0] == r.correlation == 0.0
    assert cs.detect_dual_batch(torch.zeros((0, 120000), device="cuda")).shape[0] == 0


def test_slab_transform_path_matches_the_default_path():
    """RIA_CHIRP_SLAB=1 (stages 2+3, products and inverse 3+2 in one kernel) is read once per process, so
    it runs in a child: detections must equal the default staged path on the same seeded windows."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json; sys.path.insert(0, %r)\n"
        "import numpy as np, torch, ria_b200\n"
        "from ria_b200 import sync\n"
        "pre = sync.chirp_generate_host()\n"
        "rng = np.random.default_rng(3)\n"
        "rows = []\n"
        "for i in range(12):\n"
        "    x = rng.standard_normal(120000).astype(np.float32) * np.float32(0.3)\n"
        "    p = 2000 + 5000 * i\n"
        "    x[p:p + len(pre)] += pre\n"
        "    rows.append(x)\n"
        "r = sync.results(sync.ChirpSync(ctx=ria_b200.Context(0)).detect_dual_batch(torch.from_numpy(np.stack(rows)).cuda()))\n"
        "print(json.dumps([[int(a), int(b), float(c)] for a, b, c in zip(r['detected'], r['start_sample'], r['correlation'])]))\n"
    ) % root
    outs = []
    for slab in ("0", "1"):
        env = dict(os.environ, RIA_CHIRP_SLAB=slab)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        outs.append(json.loads(res.stdout.strip().splitlines()[-1]))
    assert all(d[0] == 1 for d in outs[0])
    for a, b in zip(*outs):
        assert a[0] == b[0] and a[1] == b[1]
        assert abs(a[2] - b[2]) <= 1e-5 * abs(a[2])
