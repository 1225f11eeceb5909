"""Host side of the NVLink peer-memory exchange (csrc/peer.cu, include/acg_b200.h "Data parallelism").

One `Mailbox` per process: a device segment every peer maps through a cudaIpc handle.  `new_slot()` hands out exchange
slots; every rank creates its slots in the same order, so a slot has the same byte offset in every mailbox."""
import ctypes as C
import os

import torch

from . import _lib

MAX_PEERS = 8
SEGMENT_BYTES = 32 << 20
MAX_SLOTS = 1024


def slot_bytes(cap, world):
    n = int(_lib.load().acg_peer_slot_bytes(int(cap), int(world)))
    if n < 0:
        raise RuntimeError("acg_peer_slot_bytes: invalid arguments (cap %d, world %d)" % (cap, world))
    return n


class SlotPlan:
    """Byte offsets of the slots of one mailbox (pure host logic: same sequence of requests -> same offsets)."""

    def __init__(self, world, segment_bytes=SEGMENT_BYTES, slot_bytes_fn=slot_bytes):
        self.world, self.segment_bytes, self.next_off, self.count = world, segment_bytes, 0, 0
        self._slot_bytes = slot_bytes_fn

    def take(self, cap):
        off, idx = self.next_off, self.count
        self.next_off += self._slot_bytes(cap, self.world)
        self.count += 1
        if self.next_off > self.segment_bytes or self.count > MAX_SLOTS:
            raise RuntimeError("peer mailbox exhausted (%d bytes, %d slots)" % (self.next_off, self.count))
        return off, idx


class Mailbox:
    def __init__(self, dist, group, device, timeout_s=None):
        """timeout_s: how long an exchange kernel spins for a peer before it traps (a lost rank must not hang the GPU).
        Default 120 s, ACG_PEER_TIMEOUT_S overrides: rank skew beyond that (a rank-0-only checkpoint, a slow loader,
        a debugger) would otherwise kill a healthy job."""
        if timeout_s is None:
            timeout_s = float(os.environ.get("ACG_PEER_TIMEOUT_S", "120"))
        self._dist, self._group, self._own = dist, group, None
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > MAX_PEERS:
            raise RuntimeError("peer exchange supports at most %d ranks (one NVSwitch domain)" % MAX_PEERS)
        self.device = torch.device(device)
        self.timeout_s = float(timeout_s)
        own = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.call("acg_peer_alloc", SEGMENT_BYTES, C.byref(own))
            self._own = own
            handle = C.create_string_buffer(64)
            _lib.call("acg_peer_export", own, handle)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            self.ptrs = (C.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs[r] = own.value
                else:
                    p = C.c_void_p()
                    _lib.call("acg_peer_open", C.create_string_buffer(handles[r], 64), C.byref(p))
                    self.ptrs[r] = p.value
        self.plan = SlotPlan(self.world)
        self.epochs = torch.zeros(MAX_SLOTS, dtype=torch.int64, device=self.device)
        dist.barrier(group=group)       # every segment is mapped everywhere before the first push

    def close(self):
        """Unmap the peers' segments and free the own one (after a barrier: nobody may still push into it)."""
        if self._own is None:
            return
        own, self._own = self._own, None
        try:
            torch.cuda.synchronize(self.device)
            self._dist.barrier(group=self._group)
        except Exception:
            pass                      # process group already gone: still release the local mappings
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r != self.rank and self.ptrs[r]:
                    _lib.call("acg_peer_close", C.c_void_p(self.ptrs[r]))
            _lib.call("acg_peer_free", own)

    def new_slot(self, cap):
        off, idx = self.plan.take(cap)
        return (off, int(cap), self.epochs[idx:idx + 1])

    def exchange_struct(self, slot):
        """struct acg_peer_exchange of a slot, for launches whose last CTA does the exchange itself"""
        off, cap, epoch = slot
        return _lib.PeerExchange(C.cast(self.ptrs, C.c_void_p), _lib.ptr(epoch), off, self.rank, self.world, cap,
                                 self.timeout_s)

    def allreduce_f64(self, vec, n, slot, bn=None):
        """vec[0:n] <- sum over ranks.  bn = (C, beta, global_rows, eps, mean, rstd, scale, shift) also finalises the
        batch-norm coefficients in the same launch."""
        off, cap, epoch = slot
        if vec.dtype != torch.float64:
            raise RuntimeError("peer exchange vectors are fp64")
        if bn is None:
            args = (0, None, 0, 0.0, None, None, None, None)
        else:
            Cc, beta, rows, eps, mean, rstd, scale, shift = bn
            args = (Cc, _lib.ptr(beta), rows, eps, _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(scale), _lib.ptr(shift))
        _lib.call("acg_peer_allreduce_f64", _lib.ptr(vec), n, cap, off, self.rank, self.world, self.ptrs,
                  _lib.ptr(epoch), self.timeout_s, *args, _lib.stream())
