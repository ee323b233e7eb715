"""GPU: error counters produced by a kernel (ria_frame_counters_dev) and their NCCL reduction through the C ABI
(ria_counters_allreduce), SURVEY.md 8(b) / 8(e)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _random_status(n, seed):
    from ria_b200 import ofdm
    rng = np.random.default_rng(seed)
    st = np.zeros(n, ofdm.FRAME_STATUS_DTYPE)
    st["cw_ok"] = rng.random((n, 4)) > 0.1
    st["all_ok"] = st["cw_ok"].all(axis=1)
    st["header_valid"] = st["cw_ok"][:, 0] & (rng.random(n) > 0.05)
    st["frame_crc_ok"] = st["all_ok"] & st["header_valid"] & (rng.random(n) > 0.03)
    return st


def _expected(st):
    ok = (st["all_ok"] == 1) & (st["header_valid"] == 1) & (st["frame_crc_ok"] == 1)
    return np.array([len(st), ok.sum(), 4 * len(st), (st["cw_ok"] == 0).sum(), 0, 0, (st["header_valid"] != 1).sum(),
                     ((st["all_ok"] == 1) & ~ok).sum()], np.int64)


def test_frame_counters_kernel_matches_numpy(ctx):
    import torch
    from ria_b200 import dist as rdist
    for n in (1, 31, 1000, 70001):
        st = _random_status(n, n)
        dev = torch.from_numpy(st.view(np.uint8).reshape(n, -1)).cuda()
        got = rdist.frame_counters_dev(dev, ctx)
        got = rdist.frame_counters_dev(dev, ctx, got)            # accumulates
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), 2 * _expected(st)), n
        assert np.array_equal(rdist.frame_counters(dev).cpu().numpy()[:4], _expected(st)[:4])


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import ria_b200
from ria_b200 import dist as rdist
from tests.test_counters_gpu import _random_status, _expected
rank, world, local = rdist.env()
torch.cuda.set_device(local)
dist.init_process_group("gloo", rank=rank, world_size=world)          # only carries the 128-byte NCCL id
ctx = ria_b200.Context(local)
comm = rdist.CounterComm(ctx, rank, world)
st = _random_status(5000 + rank, 40 + rank)
dev = torch.from_numpy(st.view(np.uint8).reshape(len(st), -1)).cuda()
c = rdist.frame_counters_dev(dev, ctx)
comm.allreduce(c)
torch.cuda.synchronize()
want = sum(_expected(_random_status(5000 + r, 40 + r)) for r in range(world))
assert np.array_equal(c.cpu().numpy(), want), (rank, c.cpu().numpy(), want)
comm.close(); ctx.close()
dist.destroy_process_group()
if rank == 0: print("COUNTERS_ALLREDUCE_OK", world)
'''


def test_counters_allreduce_over_nccl(tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "COUNTERS_ALLREDUCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
