"""Host-side mirror of the reference FEC interfaces over the C ABI.

Same names, argument meaning and error behaviour as the reference classes so parity tests read
like the reference's own:

  * ``LDPCDecoder``  <- ultra::LDPCDecoder (include/ultra/fec.hpp:48-81,
                        src/fec/ldpc_decoder.cpp:263-455)
  * ``LDPCCodec``    <- ultra::fec::LDPCCodec / ICodec (src/fec/ldpc_codec.hpp:38-105,
                        src/fec/ldpc_codec.cpp:47-129, src/fec/codec_interface.hpp:29-98)

plus the batched entry point the GPU exists for (``decode_batch``).  Device memory and streams
come from torch; all arithmetic happens in libria_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import Context, RiaError, lib

LDPC_N = 648

# ultra::CodeRate (include/ultra/types.hpp:91-100)
R1_4, R1_3, R1_2, R2_3, R3_4, R5_6, R7_8 = range(7)


def code_params(rate: int) -> Tuple[int, int, int]:
    k, m, e = C.c_int(), C.c_int(), C.c_int()
    if lib().ria_ldpc_params(int(rate), C.byref(k), C.byref(m), C.byref(e)) != 0:
        raise ValueError(f"bad code rate {rate}")
    return k.value, m.value, e.value


def get_matrix(rate: int):
    """(row_ptr[m+1], edge_var[E]) of the H the library generated (H_rows order)."""
    k, m, e = code_params(rate)
    row_ptr = np.zeros(m + 1, dtype=np.int32)
    edge_var = np.zeros(e, dtype=np.int32)
    rc = lib().ria_ldpc_get_matrix(int(rate), row_ptr.ctypes.data, edge_var.ctypes.data)
    if rc != 0:
        raise ValueError(f"bad code rate {rate}")
    return row_ptr, edge_var


def recommended_iterations(rate: int) -> int:
    """LDPCCodec::getRecommendedIterations (src/fec/ldpc_codec.hpp:86-96)."""
    return {R3_4: 60, R2_3: 70, R1_2: 80, R1_3: 60, R1_4: 50}.get(int(rate), 50)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        if not torch.cuda.is_available():
            raise RiaError("no CUDA device: ria_b200 has no CPU fallback")
        _default_ctx = Context(torch.cuda.current_device())
    return _default_ctx


class LDPCDecoder:
    """ultra::LDPCDecoder drop-in (batch = 1 semantics) + ``decode_batch``."""

    def __init__(self, rate: int, ctx: Optional[Context] = None):
        code_params(rate)
        self._rate = int(rate)
        self._ctx = ctx
        self._max_iter = 50          # Impl::max_iterations default, ldpc_decoder.cpp:43
        self._factor = 0.75          # Impl::min_sum_factor default, ldpc_decoder.cpp:44
        self._last_success = False
        self._last_iters = 0

    # ---- reference setters / getters (ldpc_decoder.cpp:431-455) ----
    def setRate(self, rate: int) -> None:
        code_params(rate)
        self._rate = int(rate)

    def getRate(self) -> int:
        return self._rate

    def setMaxIterations(self, max_iter: int) -> None:
        self._max_iter = int(max_iter)

    def setMinSumFactor(self, factor: float) -> None:
        self._factor = float(factor)

    def lastDecodeSuccess(self) -> bool:
        return self._last_success

    def lastIterations(self) -> int:
        return self._last_iters

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    # ---- batched device entry point ----
    def decode_batch(self, llr: torch.Tensor, info_stride: Optional[int] = None):
        """llr: CUDA fp32 [n_cw, 648] -> (info u8 [n_cw, stride], ok u8 [n_cw], iters i32 [n_cw])."""
        if not (isinstance(llr, torch.Tensor) and llr.is_cuda):
            raise RiaError("decode_batch wants a CUDA tensor (no CPU fallback); use decode_batch_host "
                           "for host buffers")
        if llr.dtype != torch.float32 or llr.dim() != 2 or llr.shape[1] != LDPC_N:
            raise ValueError("llr must be fp32 [n_cw, 648]")
        llr = llr.contiguous()
        k, _, _ = code_params(self._rate)
        stride = int(info_stride) if info_stride else (k + 7) // 8
        n = llr.shape[0]
        info = torch.empty((n, stride), dtype=torch.uint8, device=llr.device)
        ok = torch.empty((n,), dtype=torch.uint8, device=llr.device)
        iters = torch.empty((n,), dtype=torch.int32, device=llr.device)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(llr.device))
        ctx.check(lib().ria_ldpc_decode_batch_dev(
            ctx.handle, self._rate, self._max_iter, self._factor, llr.data_ptr(), n,
            info.data_ptr(), stride, ok.data_ptr(), iters.data_ptr()))
        return info, ok, iters

    def robust_decode_batch(self, llr: torch.Tensor, info_stride: Optional[int] = None):
        """Batched robustDecodeSingleCW (src/gui/modem/streaming_decoder.cpp:1028-1058): factor 0.9375 and
        the recommended iterations, then the factors 0.875/0.75/0.625/0.5 on the codewords that failed.
        -> (info, ok, iters, attempt); attempt 0 = first decode, 1..4 = retry, 255 = all failed."""
        if not (isinstance(llr, torch.Tensor) and llr.is_cuda):
            raise RiaError("robust_decode_batch wants a CUDA tensor (no CPU fallback)")
        if llr.dtype != torch.float32 or llr.dim() != 2 or llr.shape[1] != LDPC_N:
            raise ValueError("llr must be fp32 [n_cw, 648]")
        llr = llr.contiguous()
        k, _, _ = code_params(self._rate)
        stride = int(info_stride) if info_stride else (k + 7) // 8
        n = llr.shape[0]
        info = torch.empty((n, stride), dtype=torch.uint8, device=llr.device)
        ok = torch.empty((n,), dtype=torch.uint8, device=llr.device)
        iters = torch.empty((n,), dtype=torch.int32, device=llr.device)
        attempt = torch.empty((n,), dtype=torch.uint8, device=llr.device)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(llr.device))
        ctx.check(lib().ria_ldpc_robust_decode_batch_dev(
            ctx.handle, self._rate, llr.data_ptr(), n, info.data_ptr(), stride, ok.data_ptr(), iters.data_ptr(),
            attempt.data_ptr()))
        return info, ok, iters, attempt

    def decode_batch_host(self, llr: np.ndarray, info_stride: Optional[int] = None):
        """Host buffers in, host buffers out (H2D/D2H inside the C call)."""
        llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, LDPC_N)
        k, _, _ = code_params(self._rate)
        stride = int(info_stride) if info_stride else (k + 7) // 8
        n = llr.shape[0]
        info = np.empty((n, stride), dtype=np.uint8)
        ok = np.empty((n,), dtype=np.uint8)
        iters = np.empty((n,), dtype=np.int32)
        ctx = self.ctx
        ctx.check(lib().ria_ldpc_decode_batch_host(
            ctx.handle, self._rate, self._max_iter, self._factor, llr.ctypes.data, n,
            info.ctypes.data, stride, ok.ctypes.data, iters.ctypes.data))
        return info, ok, iters

    # ---- reference single-call semantics (ldpc_decoder.cpp:284-429) ----
    def decodeSoft(self, llrs) -> bytes:
        llrs = np.asarray(llrs, dtype=np.float32).ravel()
        if llrs.size == 0:                       # :286-289
            self._last_success = False
            return b""
        k, _, _ = code_params(self._rate)
        n_full = llrs.size // LDPC_N
        rem = llrs.size - n_full * LDPC_N
        single = llrs.size <= LDPC_N             # :297-300
        n_blocks = 1 if single else n_full + (1 if rem else 0)
        padded = np.zeros((n_blocks, LDPC_N), dtype=np.float32)
        padded.ravel()[: llrs.size] = llrs
        info, ok, iters = self.decode_batch_host(padded)
        self._last_iters = int(iters[-1])
        if single:
            self._last_success = bool(ok[0])
            return info[0, : (k + 7) // 8].tobytes()
        # multi-block: bit-level concatenation of the k info bits of each block (:302-428).
        # last_success: AND over whole blocks, but a trailing partial block goes through
        # decodeBP which overwrites the flag with its own result (:401).
        self._last_success = bool(ok[-1]) if rem else bool(ok[:n_full].all())
        bits = np.unpackbits(info[:, : (k + 7) // 8], axis=1)[:, :k].ravel()
        return np.packbits(bits).tobytes()

    def decode(self, coded: bytes) -> bytes:
        """Hard-bit entry (ldpc_decoder.cpp:268-282): bit -> -6/+6 LLR."""
        bits = np.unpackbits(np.frombuffer(bytes(coded), dtype=np.uint8))
        return self.decodeSoft(np.where(bits == 1, -6.0, 6.0).astype(np.float32))


def burst_deinterleave_batch(physical: torch.Tensor, ctx: Optional[Context] = None) -> torch.Tensor:
    """fec::BurstInterleaver::deinterleave (src/fec/burst_interleaver.cpp:39-78) for a batch of burst
    groups: physical CUDA fp32 [n_groups, N, >= 2592] -> logical [n_groups, N, 2592]."""
    if not (isinstance(physical, torch.Tensor) and physical.is_cuda and physical.dtype == torch.float32 and physical.dim() == 3):
        raise ValueError("physical must be a CUDA fp32 [n_groups, N, >= 2592] tensor")
    if physical.shape[2] < 2592:
        raise ValueError("BurstInterleaver::deinterleave: soft bits size mismatch")
    physical = physical.contiguous()
    g, n, stride = physical.shape
    out = torch.empty((g, n, 2592), dtype=torch.float32, device=physical.device)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(physical.device))
    ctx.check(lib().ria_burst_deinterleave_batch_dev(ctx.handle, physical.data_ptr(), stride, n, g, out.data_ptr(), 2592))
    return out


def ladder_perturb_batch(llr: torch.Tensor, attempt: int, ctx: Optional[Context] = None) -> torch.Tensor:
    """Soft bits that attempt `attempt` (1..38) of decodeFixedFrame's retry ladder decodes
    (src/protocol/frame_v2.cpp:1409-1542).  llr: CUDA fp32 [n_cw, 648]."""
    if not (isinstance(llr, torch.Tensor) and llr.is_cuda and llr.dtype == torch.float32 and llr.dim() == 2
            and llr.shape[1] == LDPC_N):
        raise ValueError("llr must be a CUDA fp32 [n_cw, 648] tensor")
    llr = llr.contiguous()
    out = torch.empty_like(llr)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(llr.device))
    ctx.check(lib().ria_ldpc_ladder_perturb_dev(ctx.handle, llr.data_ptr(), llr.shape[0], int(attempt), out.data_ptr()))
    return out


@dataclass
class DecodeResult:
    """fec::DecodeResult (src/fec/codec_interface.hpp:29-34)."""
    success: bool = False
    data: bytes = b""
    iterations: int = 0
    ber_estimate: float = 0.0


class LDPCCodec:
    """fec::LDPCCodec / ICodec decode side (src/fec/ldpc_codec.cpp:47-129)."""

    CODEWORD_BITS = 648
    CODEWORD_BYTES = 81

    def __init__(self, rate: int = R1_2, ctx: Optional[Context] = None):
        self._rate = int(rate)
        self._max_iterations = recommended_iterations(rate)
        self._decoder = LDPCDecoder(rate, ctx)
        self._decoder.setMaxIterations(self._max_iterations)

    def getName(self) -> str:
        return "802.11n LDPC"

    def setRate(self, rate: int) -> None:
        self._rate = int(rate)
        self._decoder.setRate(rate)
        rec = recommended_iterations(rate)
        if self._max_iterations != rec:
            self._max_iterations = rec
            self._decoder.setMaxIterations(rec)

    def getRate(self) -> int:
        return self._rate

    def setMaxIterations(self, iterations: int) -> None:
        self._max_iterations = int(iterations)
        self._decoder.setMaxIterations(iterations)

    def getMaxIterations(self) -> int:
        return self._max_iterations

    def decode(self, soft_bits) -> Tuple[bool, bytes]:
        data = self._decoder.decodeSoft(soft_bits)
        return self._decoder.lastDecodeSuccess(), data

    def decodeExtended(self, soft_bits) -> DecodeResult:
        r = DecodeResult()
        r.data = self._decoder.decodeSoft(soft_bits)
        r.success = self._decoder.lastDecodeSuccess()
        r.iterations = self._decoder.lastIterations()
        r.ber_estimate = (np.float32(r.iterations) / np.float32(self._max_iterations * 10.0)
                          if r.success else 0.5)
        return r

    def getCodewordBits(self) -> int:
        return self.CODEWORD_BITS

    def getInfoBits(self) -> int:
        return code_params(self._rate)[0]

    def getParityBits(self) -> int:
        return self.CODEWORD_BITS - self.getInfoBits()

    def getCodewordBytes(self) -> int:
        return self.CODEWORD_BYTES

    def getDataBytes(self) -> int:
        return self.getInfoBits() // 8

    def getEffectiveRate(self) -> float:
        return self.getInfoBits() / self.CODEWORD_BITS

    @property
    def decoder(self) -> LDPCDecoder:
        return self._decoder


class ChaseCache:
    """fec::ChaseCache (src/fec/chase_cache.{hpp,cpp}) with the soft bits resident in HBM.

    The policy is the reference's host logic: key = (seq, src_hash, dst_hash), at most
    MAX_COMBINES (4) receptions per codeword, ``max_entries`` entries with LRU eviction, TTL.
    ``store_batch`` performs one reception round for many (key, cw_index) items with one kernel
    (ria_chase_combine_batch_dev); a (key, cw) pair may appear once per round."""

    MAX_COMBINES = 4
    LDPC_BLOCK_SIZE = 648

    def __init__(self, max_entries: int = 16, entry_ttl_s: float = 30.0, max_cw: int = 16,
                 ctx: Optional[Context] = None, device=None):
        import time as _time
        self._time = _time
        self.max_entries, self.ttl, self.max_cw = int(max_entries), float(entry_ttl_s), int(max_cw)
        self._ctx = ctx
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.acc = torch.zeros((self.max_entries * self.max_cw, LDPC_N), dtype=torch.float32, device=self.device)
        self.entries = {}          # key -> dict(slot, total_cw, counts[], decoded[], created, last)
        self.free = list(range(self.max_entries))
        self.enabled = True
        self.stats = dict(stores=0, combines=0, cache_hits=0, cache_misses=0, entries_evicted=0,
                          entries_expired=0, recoveries=0)

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def _release(self, key, stat, busy, flush):
        """Drop an entry and recycle its HBM rows.  If an item of the current round already targets those
        rows, the pending combine launch goes out first: the kernel's contract is one writer per row and
        launch, and the sequential reference semantics are store-then-evict."""
        slot = self.entries.pop(key)["slot"]
        if slot in busy:
            flush()
        self.free.append(slot)
        if stat:
            self.stats[stat] += 1

    def _prune(self, busy=(), flush=lambda: None):
        now = self._time.monotonic()
        for k in [k for k, e in self.entries.items() if now - e["created"] > self.ttl]:
            self._release(k, "entries_expired", busy, flush)

    def _evict(self, busy=(), flush=lambda: None):
        while len(self.entries) >= self.max_entries:
            k = min(self.entries, key=lambda q: self.entries[q]["last"])
            self._release(k, "entries_evicted", busy, flush)

    def _launch(self, items, slots, first, soft):
        if not items:
            return
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(self.device))
        idx = np.asarray(items, np.int64)
        d_slots = torch.from_numpy(slots[idx]).to(self.device)
        d_first = torch.from_numpy(first[idx]).to(self.device)
        rows = soft if len(idx) == soft.shape[0] else soft.index_select(0, torch.from_numpy(idx).to(self.device))
        ctx.check(lib().ria_chase_combine_batch_dev(
            ctx.handle, self.acc.data_ptr(), d_slots.data_ptr(), d_first.data_ptr(), rows.data_ptr(),
            rows.stride(0), len(idx)))

    def store_batch(self, keys, cw_indices, total_cws, soft: torch.Tensor):
        """One reception round: item i stores soft[i] (CUDA fp32 [n, >=648]) under keys[i] /
        cw_indices[i].  Returns a list of bools like ChaseCache::store.  Items are applied in order with the
        reference's sequential semantics; normally that is ONE kernel launch, and an extra launch only when
        an eviction / expiry inside the round recycles rows that an earlier item of the round wrote."""
        n = len(keys)
        assert soft.is_cuda and soft.dtype == torch.float32 and soft.shape[0] == n and soft.shape[1] >= LDPC_N
        if soft.stride(1) != 1:
            soft = soft.contiguous()
        slots = np.full(n, -1, np.int32)
        first = np.zeros(n, np.uint8)
        ok = [False] * n
        seen = set()
        pending, busy = [], set()          # items not yet launched / entry slots they write

        def flush():
            self._launch(pending, slots, first, soft)
            pending.clear()
            busy.clear()

        for i, (key, cw, total) in enumerate(zip(keys, cw_indices, total_cws)):
            if not self.enabled or cw < 0 or cw >= total or total <= 0 or total > self.max_cw or (key, cw) in seen:
                continue
            self.stats["stores"] += 1
            self._prune(busy, flush)
            e = self.entries.get(key)
            now = self._time.monotonic()
            if e is None:
                self._evict(busy, flush)
                e = dict(slot=self.free.pop(), total_cw=total, counts=[0] * total, decoded=[False] * total,
                         created=now, last=now)
                self.entries[key] = e
            e["last"] = now
            if cw >= e["total_cw"] or e["decoded"][cw] or e["counts"][cw] >= self.MAX_COMBINES:
                continue
            slots[i] = e["slot"] * self.max_cw + cw
            first[i] = 1 if e["counts"][cw] == 0 else 0
            if e["counts"][cw] > 0:
                self.stats["combines"] += 1
            e["counts"][cw] += 1
            seen.add((key, cw))
            pending.append(i)
            busy.add(e["slot"])
            ok[i] = True
        flush()
        return ok

    def store(self, key, cw_index: int, soft_bits, total_cw: int) -> bool:
        soft = torch.as_tensor(np.asarray(soft_bits, np.float32)).to(self.device).reshape(1, -1)
        if soft.shape[1] != LDPC_N:
            return False
        return self.store_batch([key], [cw_index], [total_cw], soft)[0]

    def getCombined(self, key, cw_index: int):
        e = self.entries.get(key)
        if (not self.enabled or e is None or cw_index < 0 or cw_index >= e["total_cw"]
                or e["counts"][cw_index] == 0 or e["decoded"][cw_index]):
            self.stats["cache_misses"] += 1
            return None
        self.stats["cache_hits"] += 1
        return self.acc[e["slot"] * self.max_cw + cw_index]

    def getCombineCount(self, key, cw_index: int) -> int:
        e = self.entries.get(key)
        return e["counts"][cw_index] if e is not None and 0 <= cw_index < e["total_cw"] else 0

    def markDecoded(self, key, cw_index: int) -> None:
        e = self.entries.get(key)
        if e is None:
            return
        if 0 <= cw_index < e["total_cw"]:
            e["decoded"][cw_index] = True
        if all(e["decoded"]):
            self.free.append(self.entries.pop(key)["slot"])

    def removeEntry(self, key) -> None:
        """ChaseCache::removeEntry: drop the entry of a frame that decoded completely."""
        e = self.entries.pop(key, None)
        if e is not None:
            self.free.append(e["slot"])

    def size(self) -> int:
        return len(self.entries)

    def clear(self) -> None:
        for e in self.entries.values():
            self.free.append(e["slot"])
        self.entries.clear()
