"""Peer-memory exchange between the ranks of one node (host side of csrc/peer.cu).

Every rank allocates a small mailbox, shares it through CUDA IPC (the handles travel once through
``dist.all_gather_object``) and maps the other ranks' mailboxes; ``allreduce_`` then is ONE single-CTA kernel per rank
that writes into all mailboxes over NVLink / NVSwitch and spins on its own.  Used for the four BatchNorm statistic
exchanges per step of ``naiveSyncBN1d`` (mmdet3d/ops/norm.py:28-86), which are latency-, not bandwidth-bound.

``PeerExchange.get(device)`` returns the process-wide instance, or None when the ranks are not all on one node, peer
access is not available or the IPC mapping fails — the caller then stays on ``torch.distributed`` (and a warning
says so once)."""
from __future__ import annotations

import ctypes as C
import os
import warnings

import torch
import torch.distributed as dist

from . import lib as L

_INSTANCE = {}


class PeerExchange:
    def __init__(self, device):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 8:
            raise RuntimeError("more than 8 ranks")
        dev = torch.device(device)
        self.device = dev
        self.timeout = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ptrs = [None] * self.world
        with torch.cuda.device(dev):
            own, handle = C.c_void_p(), C.create_string_buffer(64)
            L.run("peer_mailbox_create", self.world, C.byref(own), handle)
            self.ptrs[self.rank] = own.value
            everyone = [None] * self.world          # (hostname, device index, 64-byte IPC handle) of every rank
            dist.all_gather_object(everyone, (os.uname().nodename, dev.index, handle.raw))
            problem = None
            try:
                if len({e[0] for e in everyone}) != 1:
                    raise RuntimeError("ranks span several nodes")
                for r, (_, peer_dev, raw) in enumerate(everyone):
                    if r == self.rank:
                        continue
                    if not torch.cuda.can_device_access_peer(dev.index, peer_dev):
                        raise RuntimeError(f"device {dev.index} cannot access device {peer_dev}")
                    mapped = C.c_void_p()
                    L.run("peer_mailbox_open", C.create_string_buffer(raw, 64), C.byref(mapped))
                    self.ptrs[r] = mapped.value
            except Exception as e:      # noqa: BLE001
                problem = e
        # all ranks or none — and nobody writes into a mailbox that is not mapped everywhere yet (acts as the barrier)
        flag = torch.tensor([0 if problem else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise RuntimeError(str(problem) if problem else "another rank could not map the mailboxes")
        self.ctx = L.PeerCtx()
        self.ctx.rank, self.ctx.world = self.rank, self.world
        for r, p in enumerate(self.ptrs):
            self.ctx.mailbox[r] = p
        self.ctx.timeout_flag = self.timeout.data_ptr()
        self.epoch = 0

    def allreduce_(self, buf: torch.Tensor, pre_scale=1.0, post_scale=1.0):
        """In place: buf = post_scale * sum_r(pre_scale_r * buf_r); float64, at most 512 elements, every rank calls."""
        assert buf.dtype == torch.float64 and buf.is_contiguous()
        self.epoch += 1
        L.run("peer_allreduce_f64", C.byref(self.ctx), L.ptr(buf), buf.numel(), float(pre_scale), float(post_scale),
              C.c_uint64(self.epoch), L.stream_ptr(buf.device))
        return buf

    @classmethod
    def get(cls, device, name="bn"):
        """The process-wide exchange ``name`` for ``device`` (every rank must call at the same point the first time).
        Exchanges that may be in flight at the same time (different streams) need different names: one mailbox serves
        ONE ordered sequence of calls."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        key = (torch.device(device).index, name)
        if key not in _INSTANCE:
            inst = None
            if not os.environ.get("GEOMAE_NO_PEER_EXCHANGE"):       # (set on every rank or on none)
                try:
                    inst = cls(device)
                except Exception as e:          # noqa: BLE001 — any failure means "use the collective library"
                    warnings.warn(f"geomae_b200: peer-memory exchange unavailable ({e}); BatchNorm statistics go "
                                  f"through torch.distributed")
            _INSTANCE[key] = inst
        return _INSTANCE[key]


class _RawCuda:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr, False), version=3)


class SharedGradients:
    """The flat gradient buffer of every rank, mapped everywhere, and the peer-memory form of the DDP gradient exchange
    (csrc/peer.cu: k_peer_reduce_shard between two barrier exchanges).  Collective constructor; raises when the node
    cannot do it (the trainer then keeps NCCL)."""

    def __init__(self, device, n_floats):
        dev = torch.device(device)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.early, self.late = PeerExchange.get(dev, "grad_early"), PeerExchange.get(dev, "grad_late")
        if self.early is None or self.late is None:
            raise RuntimeError("peer mailboxes unavailable")
        ptrs, problem = [None] * self.world, None
        with torch.cuda.device(dev):
            own, handle = C.c_void_p(), C.create_string_buffer(64)
            L.run("peer_buffer_create", 4 * n_floats, C.byref(own), handle)
            ptrs[self.rank] = own.value
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handle.raw)
            try:
                for r, raw in enumerate(everyone):
                    if r != self.rank:
                        mapped = C.c_void_p()
                        L.run("peer_mailbox_open", C.create_string_buffer(raw, 64), C.byref(mapped))
                        ptrs[r] = mapped.value
            except Exception as e:      # noqa: BLE001
                problem = e
        flag = torch.tensor([0 if problem else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise RuntimeError(str(problem) if problem else "another rank could not map the gradient buffers")
        self.ptrs = (C.c_void_p * 8)(*ptrs)
        self.flat_grad = torch.as_tensor(_RawCuda(own.value, n_floats, "<f4"), device=dev)
        self.token = torch.zeros(2, dtype=torch.float64, device=dev)
        self.stream = torch.cuda.Stream(dev)

    def exchange(self, px: PeerExchange, ranges):
        """Sum the element ranges over the ranks, in place in every rank's buffer, on the CURRENT stream."""
        tok = self.token[:1] if px is self.early else self.token[1:]
        px.allreduce_(tok)                          # barrier: every rank's gradients of these ranges are complete
        s = L.stream_ptr(self.flat_grad.device)
        for a, b in ranges:
            if b > a:
                L.run("peer_reduce_shard", C.byref(px.ctx), self.ptrs, a, b, s)
        px.allreduce_(tok)                          # barrier: every slice has been written into every buffer


def scale_(buf: torch.Tensor, scale: float):
    """Single-rank form of the same kernel: buf *= scale in one launch of the library (no tensor-library launch)."""
    ctx = L.PeerCtx()
    ctx.rank, ctx.world = 0, 1
    L.run("peer_allreduce_f64", C.byref(ctx), L.ptr(buf), buf.numel(), float(scale), 1.0, C.c_uint64(0),
          L.stream_ptr(buf.device))
    return buf
