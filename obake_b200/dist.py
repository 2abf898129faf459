"""Multi-GPU series product: one process per GPU, torch.distributed for the plumbing.

The left (shorter) operand's terms are split into `world` contiguous slices; rank r multiplies slice r
by the whole right operand (dense row kernel or segment-owner scatter kernel, whichever the engine picks)
and leaves a partial product grouped by key mod 4093, a fixed prime.  Group s is owned by rank
s * world // 4093 (contiguous ranges), so the rows a rank must send to each peer are contiguous:
one all-to-all of counts, one of keys, one of coefficient limbs (NCCL over NVLink on the GPU box; gloo in
the CPU tests), then the owner sums what it received by key.  No other collective is on the data path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

import ctypes as C

from . import _capi
from .engine import Engine, Product, _raise, raw_view


class _DevArray:
    """Zero-copy torch view of engine-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def device_tensor(ptr: int, shape, device) -> torch.Tensor:
    if int(np.prod(shape)) == 0:
        return torch.empty(tuple(shape), dtype=torch.int64, device=device)
    return torch.as_tensor(_DevArray(ptr, shape), device=device)


def owner_ranges(nsegs: int, world: int):
    return [(nsegs * g // world, nsegs * (g + 1) // world) for g in range(world)]


def exchange_partials(keys: torch.Tensor, cfs: torch.Tensor, send_counts, group=None):
    """All-to-all of the rows of `keys` [n, W] and `cfs` [n, L] (int64 bit patterns); rows are already ordered
    by destination rank with `send_counts[g]` rows for rank g.  Returns (recv_keys, recv_cfs, recv_counts)."""
    world = dist.get_world_size(group)
    dev = keys.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(v) for v in rc.tolist()]
    nrecv = sum(recv_counts)
    W, L = keys.shape[1], cfs.shape[1]
    rkeys = torch.empty((nrecv, W), dtype=keys.dtype, device=dev)
    rcfs = torch.empty((nrecv, L), dtype=cfs.dtype, device=dev)
    dist.all_to_all_single(rkeys, keys.contiguous(), output_split_sizes=recv_counts,
                           input_split_sizes=list(send_counts), group=group)
    dist.all_to_all_single(rcfs, cfs.contiguous(), output_split_sizes=recv_counts,
                           input_split_sizes=list(send_counts), group=group)
    return rkeys, rcfs, recv_counts


def owner_mul(engine: Engine, vx, vy, limit=None, idx=None, group=None):
    """Zero-exchange alternative: every rank holds both operands and forms only the output segments /
    output rows it owns, so there is no partial product, no all-to-all and no merge.  Returns (Product in DEVICE
    memory with this rank's share of the result, info dict).  If a shared-memory segment table overflows on any
    rank, ALL ranks retry with a finer segmentation (they must agree on it: ownership is defined by it)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", engine.device)
    engine.set_option("mgpu_owner", 1)
    fill = 62
    try:
        for attempt in range(6):
            part, failed = None, 0
            try:
                part, counts = engine.mul_partial_views(vx, vy, rank, world, limit, idx)
            except OverflowError as ex:
                if "segment table" not in str(ex):
                    raise
                failed = 1
            # The dense kernels have no table to overflow, and every rank takes the same kernel decision (it depends on
            # the statistics of the FULL operands): no verdict to agree on, no collective at all on that path.
            dense = failed == 0 and part.stats["kernel"] in (1, 2, 3)
            if not dense:
                flag = torch.tensor([failed], dtype=torch.int64, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if dense or int(flag.item()) == 0:
                return part, {"mode": "owner", "partial_stats": part.stats, "attempts": attempt + 1,
                              "sent_rows": 0, "recv_rows": 0, "bytes_sent": 0, "partial_terms": part.nterms,
                              "nsegs": part.stats["nsegs"]}
            if part is not None:
                part.free()
            fill = max(4, fill // 2)
            engine.set_option("fill_pct", fill)
        raise OverflowError("owner-computes product: segment tables keep overflowing")
    finally:
        engine.set_option("mgpu_owner", 0)
        engine.set_option("fill_pct", 62)


def sharded_mul(engine: Engine, vx, vy, limit=None, idx=None, group=None, result_mem=_capi.OBK_MEM_DEVICE):
    """This rank's share of x*y: partial product -> all-to-all by owner -> merge.  Returns (Product holding the
    segments this rank owns, info dict)."""
    import os
    import time
    timing = os.environ.get("OBK_DIST_TIMING") == "1"
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", engine.device)
    t0 = time.perf_counter()
    part, counts = engine.mul_partial_views(vx, vy, rank, world, limit, idx)
    t1 = time.perf_counter()
    nsegs = _capi.OBK_EXCHANGE_MOD  # fixed modulus: owner ranges are the same on every rank
    W, L = part.key_words, part.cf_limbs
    pk = device_tensor(part.keys_ptr, (part.nterms, W), dev)
    pc = device_tensor(part.cfs_ptr, (part.nterms, L), dev)
    rkeys, rcfs, recv_counts = exchange_partials(pk, pc, counts, group)
    info = {"partial_terms": part.nterms, "sent_rows": sum(counts) - counts[rank], "recv_rows": sum(recv_counts),
            "bytes_sent": (sum(counts) - counts[rank]) * 8 * (W + L), "partial_stats": part.stats,
            "nsegs": nsegs}
    is_f64 = part.is_f64
    torch.cuda.current_stream(dev).synchronize()
    t2 = time.perf_counter()
    view = raw_view(rkeys.shape[0], W, vx.key_signed, vx.nvars, vx.psize,
                    _capi.OBK_CF_F64 if is_f64 else _capi.OBK_CF_INT, L, _capi.OBK_MEM_DEVICE,
                    rkeys.data_ptr() if rkeys.numel() else 0, rcfs.data_ptr() if rcfs.numel() else 0,
                    vx.key_bits or 64)
    merged = engine.merge_view(view, nsegs, result_mem)
    info["merge_stats"] = merged.stats
    if timing:
        t3 = time.perf_counter()
        info["timing_ms"] = {"partial": (t1 - t0) * 1e3, "exchange": (t2 - t1) * 1e3, "merge": (t3 - t2) * 1e3,
                             "partial_device_ms": part.stats["ms_total"], "merge_device_ms": merged.stats["ms_total"]}
    part.free()
    return merged, info


class PeerExchange:
    """Receive windows in NVLink peer memory for the fused push exchange (include/obk_b200.h, obk_peer_windows).

    Every rank allocates one window (2 parities of `window_bytes`), exports it with CUDA IPC and opens the windows
    of its peers; the 64-byte handles travel through torch.distributed (plumbing).  `pushed_mul` then replaces
    regroup-by-owner + all-to-all of counts + two NCCL all-to-alls by P2P stores issued by the producing rank, with
    ONE one-word all-reduce as the stream-ordered barrier between the pushes and the owner-side merge."""

    def __init__(self, engine: Engine, window_bytes: int = 1 << 30, group=None):
        self.engine, self.group = engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.window_bytes = int(window_bytes)
        self.parity = 0
        L = _capi.lib()
        # The all-reduce between the pushes and the merge is only a barrier for the engine's work if both run on
        # the SAME stream: bind the engine to torch's current stream (NCCL collectives are ordered on it).
        engine.set_stream(torch.cuda.current_stream(torch.device("cuda", engine.device)).cuda_stream)
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        rc = L.obk_peer_window_alloc(engine._h, self.window_bytes, C.byref(own), handle)
        if rc != 0:
            _raise(rc, engine.last_error())
        self._own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._base = (C.c_void_p * self.world)()
        self._opened = []
        for g in range(self.world):
            if g == self.rank:
                self._base[g] = own.value
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[g])
            rc = L.obk_peer_window_open(engine._h, buf, C.byref(p))
            if rc != 0:
                _raise(rc, engine.last_error())
            self._base[g] = p.value
            self._opened.append(p)
        dev = torch.device("cuda", engine.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.barrier(group=group)

    def _windows(self) -> _capi.PeerWindows:
        return _capi.PeerWindows(self.world, self.rank, self.parity, 0, self.window_bytes,
                                 C.cast(self._base, C.POINTER(C.c_void_p)))

    def pushed_mul(self, vx, vy, limit=None, idx=None, result_mem=_capi.OBK_MEM_DEVICE):
        """This rank's share of x*y: partial product pushed to the owners' windows -> barrier -> merge."""
        import os
        import time
        timing = os.environ.get("OBK_DIST_TIMING") == "1"
        eng, L = self.engine, _capi.lib()
        keep = []
        tr = eng._trunc(limit, idx, keep)
        w = self._windows()
        st = _capi.Stats()
        t0 = time.perf_counter()
        rc = L.obk_poly_mul_push(eng._h, C.byref(vx), C.byref(vy), C.byref(tr) if tr is not None else None,
                                 C.byref(w), C.byref(st))
        t1 = time.perf_counter()
        failed = 0
        if rc != 0:
            failed, msg = rc, eng.last_error()
        # the barrier: stream-ordered, carries the failure flag so that every rank takes the same branch
        self._flag.fill_(failed)
        dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)
        if timing:
            torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        res, mst = _capi.PolyResult(), _capi.Stats()
        # every rank computes the accumulator width from the FULL operands (also a rank whose slice is empty)
        acc = int(st.acc_limbs) if rc == 0 and st.acc_limbs else 1
        rc2 = 0
        if rc == 0:
            rc2 = L.obk_peer_merge(eng._h, C.byref(w), C.byref(vx), acc, int(result_mem), C.byref(res), C.byref(mst))
        t3 = time.perf_counter()
        self.parity ^= 1
        self._last_timing = {"push": (t1 - t0) * 1e3, "barrier": (t2 - t1) * 1e3, "merge": (t3 - t2) * 1e3,
                             "partial_device_ms": float(st.ms_total), "merge_device_ms": float(mst.ms_total)} if timing else None
        if int(self._flag.item()) != 0 or rc != 0:
            if rc == 0 and rc2 == 0 and res.owner:
                L.obk_result_free(eng._h, C.byref(res))
            # a rank whose own push failed has not merged: whatever the peers already pushed into its window must not
            # survive until this parity is reused (obk_peer_merge is the only other place that clears the header).
            # Stream-ordered behind the all-reduce, i.e. after every peer's push of this step has landed.
            if rc != 0:
                L.obk_peer_window_reset(eng._h, C.byref(w))
            _raise(rc or int(self._flag.item()), msg if rc != 0 else "a peer failed in the pushed product")
        if rc2 != 0:
            _raise(rc2, eng.last_error())
        merged = Product(eng, res, mst)
        info = {"mode": "push", "partial_stats": st.as_dict(), "merge_stats": merged.stats,
                "partial_terms": int(st.nterms_out), "recv_rows": int(mst.term_products), "nsegs": _capi.OBK_EXCHANGE_MOD,
                "sent_rows": None, "bytes_sent": None}
        if self._last_timing:
            info["timing_ms"] = self._last_timing
        return merged, info

    def close(self):
        L = _capi.lib()
        dist.barrier(group=self.group)
        for p in self._opened:
            L.obk_peer_window_close(self.engine._h, p, 0)
        self._opened = []
        dist.barrier(group=self.group)
        if self._own:
            L.obk_peer_window_close(self.engine._h, self._own, 1)
            self._own = None
