"""Multi-GPU plumbing: one process per GPU, frames sharded statically, one all-reduce of counters.

Frames are independent units: a batch of ``n_total`` frames (or SNR-sweep points)
is split into contiguous blocks of global frame ids, every rank synthesises, demodulates and
decodes its own block (the Philox channel is keyed by the global frame id, so results do not depend
on the world size), and the only exchange is ``all_reduce(SUM)`` over an int64 counter vector --
NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist

COUNTER_NAMES = ("frames", "frames_ok", "cw", "cw_fail", "bits", "bit_err", "sync_miss", "crc_fail")


def env() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str = "nccl", device=None) -> Tuple[int, int, int]:
    rank, world, local = env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [first, first + count) of global ids owned by ``rank``; blocks differ by at
    most one element and cover 0..n_total-1 exactly once."""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks of an int64 counter tensor (any shape)."""
    assert counters.dtype == torch.int64
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value: float, device=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def frame_counters(status, sent_ok=None) -> torch.Tensor:
    """[frames, frames_ok, cw, cw_fail, 0, 0, 0, crc_fail] from a ria_frame_status array
    (numpy structured array or the uint8 device tensor produced by the chain)."""
    import numpy as np
    from .ofdm import FRAME_STATUS_DTYPE
    if isinstance(status, torch.Tensor):
        dt = FRAME_STATUS_DTYPE
        off = {k: dt.fields[k][1] for k in dt.names}
        st = status.view(torch.uint8)
        cw_ok = st[:, off["cw_ok"]:off["cw_ok"] + 4]
        ok = (st[:, off["all_ok"]] == 1) & (st[:, off["header_valid"]] == 1) & (st[:, off["frame_crc_ok"]] == 1)
        crc_fail = (st[:, off["all_ok"]] == 1) & ~ok
        c = torch.zeros(len(COUNTER_NAMES), dtype=torch.int64, device=status.device)
        c[0] = st.shape[0]; c[1] = ok.sum(); c[2] = 4 * st.shape[0]; c[3] = (cw_ok == 0).sum(); c[7] = crc_fail.sum()
        return c
    st = np.asarray(status)
    ok = (st["all_ok"] == 1) & (st["header_valid"] == 1) & (st["frame_crc_ok"] == 1)
    return torch.tensor([len(st), int(ok.sum()), 4 * len(st), int((st["cw_ok"] == 0).sum()), 0, 0, 0,
                         int(((st["all_ok"] == 1) & ~ok).sum())], dtype=torch.int64)


def frame_counters_dev(status: torch.Tensor, ctx, out: torch.Tensor = None) -> torch.Tensor:
    """COUNTER_NAMES of a ria_frame_status device array, accumulated by a kernel on the context stream
    (ria_frame_counters_dev) into an int64[8] device tensor -- no eager tensor ops, no host sync."""
    import ctypes as C
    from ._lib import lib
    from .ofdm import FRAME_STATUS_DTYPE
    assert status.is_cuda and status.dtype == torch.uint8 and status.shape[1] == FRAME_STATUS_DTYPE.itemsize
    status = status.contiguous()
    if out is None:
        out = torch.zeros(len(COUNTER_NAMES), dtype=torch.int64, device=status.device)
    ctx.set_stream(torch.cuda.current_stream(status.device))
    ctx.check(lib().ria_frame_counters_dev(ctx.handle, C.c_void_p(status.data_ptr()), status.shape[0], C.c_void_p(out.data_ptr())))
    return out


class CounterComm:
    """The path's one collective through the C ABI: ria_counters_allreduce = ncclAllReduce(sum, int64) on the
    context stream.  The NCCL communicator is the library's own: rank 0 draws the unique id
    (ria_nccl_get_unique_id), torch.distributed only carries those 128 bytes to the other ranks."""

    def __init__(self, ctx, rank: int, world: int):
        import ctypes as C
        from ._lib import lib
        self.ctx, self.rank, self.world = ctx, rank, world
        self.comm = C.c_void_p()
        uid = (C.c_char * 128)()
        if rank == 0:
            rc = lib().ria_nccl_get_unique_id(C.addressof(uid))
            if rc != 0:
                raise RuntimeError("NCCL is not loadable in this process")
        box = [bytes(uid)]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        uid = (C.c_char * 128).from_buffer_copy(box[0])
        ctx.check(lib().ria_nccl_comm_create(ctx.handle, C.addressof(uid), rank, world, C.addressof(self.comm)))

    def allreduce(self, counters: torch.Tensor) -> torch.Tensor:
        import ctypes as C
        from ._lib import lib
        assert counters.is_cuda and counters.dtype == torch.int64 and counters.is_contiguous()
        self.ctx.set_stream(torch.cuda.current_stream(counters.device))
        self.ctx.check(lib().ria_counters_allreduce(self.ctx.handle, self.comm, C.c_void_p(counters.data_ptr()), counters.numel()))
        return counters

    def close(self):
        from ._lib import lib
        if self.comm:
            lib().ria_nccl_comm_destroy(self.comm)
            self.comm = None
