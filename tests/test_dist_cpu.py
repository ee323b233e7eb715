"""CPU: the N > 1 path's host logic with gloo, world_size 2 (and 3): shard ranges cover the batch
exactly once, the counter all-reduce equals the single-process total, max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _status(n, seed):
    from ria_b200.ofdm import FRAME_STATUS_DTYPE
    rng = np.random.default_rng(seed)
    st = np.zeros(n, FRAME_STATUS_DTYPE)
    st["cw_ok"] = rng.random((n, 4)) > 0.1
    st["all_ok"] = st["cw_ok"].all(axis=1)
    st["header_valid"] = st["cw_ok"][:, 0]
    st["frame_crc_ok"] = st["all_ok"] & (rng.random(n) > 0.02)
    return st


def _worker(rank, world, port, n_total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ria_b200 import dist as rd
    r, w, _ = rd.init("gloo")
    assert (r, w) == (rank, world)
    first, count = rd.shard_range(n_total, r, w)
    st = _status(n_total, 5)[first:first + count]            # every rank derives the same global array
    c = rd.allreduce_counters(rd.frame_counters(st))
    t = rd.max_over_ranks(1.0 + rank)
    q.put((rank, first, count, c.tolist(), t))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1001), (3, 10)])
def test_shard_and_allreduce_gloo(ria_lib, world, n_total):
    from ria_b200 import dist as rd
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, first, count, c, t in res:
        covered += list(range(first, first + count))
        assert t == float(world)                           # max over ranks of (1 + rank)
    assert covered == list(range(n_total))
    want = rd.frame_counters(_status(n_total, 5)).tolist()
    for _, _, _, c, _ in res:
        assert c == want                                   # identical on every rank, equals the serial total


def test_shard_range_properties():
    from ria_b200.dist import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in blocks) == n
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
