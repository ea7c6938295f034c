"""NVLink peer-memory exchange of BatchNorm statistics (host side of csrc/peer.cu).

The reference converts every BatchNorm to torch.nn.SyncBatchNorm (train.py:97-98), whose forward/backward call
NCCL collectives.  Here each rank owns one IPC-shared buffer; `PeerExchange.allreduce` sums a small fp32 vector
across the ranks of the node in ONE kernel launch, which (unlike a NCCL call) can live inside the captured
forward/backward CUDA graphs.  Only used when every rank of the default process group sits on the same node and
peer mapping succeeds on all of them; otherwise the engine keeps torch.distributed.all_reduce.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

HANDLE_BYTES = 64
MAX_WORLD = 8
MAX_SLOTS = 1024
SLOT_FLOATS = 4096


class PeerExchange:
    def __init__(self, device: torch.device):
        L = _lib.lib()
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.device = device
        self.timeout_s = float(os.environ.get("CRIS_B200_PEER_TIMEOUT_S", "120"))
        self.ptrs = (C.c_void_p * self.world)()
        self._own: Optional[int] = None
        ok, err = 1, ""
        handle = C.create_string_buffer(HANDLE_BYTES)
        own = C.c_void_p()
        with torch.cuda.device(device):
            if self.world > MAX_WORLD:
                ok, err = 0, f"world size {self.world} > {MAX_WORLD}"
            elif L.cris_peer_buffer_create(C.byref(own), handle) != 0:
                ok, err = 0, L.cris_last_error().decode()
            # every rank learns every handle (and whether all creations worked) through the process group's store
            infos = [None] * self.world
            dist.all_gather_object(infos, (ok, bytes(handle.raw), os.uname().nodename, err))
            if all(i[0] for i in infos) and len({i[2] for i in infos}) == 1:
                self._own = own.value
                for r, info in enumerate(infos):
                    if r == self.rank:
                        self.ptrs[r] = own.value
                        continue
                    p = C.c_void_p()
                    if L.cris_peer_buffer_open(info[1], C.byref(p)) != 0:
                        ok, err = 0, L.cris_last_error().decode()
                        break
                    self.ptrs[r] = p.value
            else:
                ok = 0
                err = err or "; ".join(str(i[3]) for i in infos if i[3]) or "ranks span several nodes"
            flags = [None] * self.world
            dist.all_gather_object(flags, ok)
        self.ok = all(flags)
        self.error = err
        if not self.ok:
            self.close()

    def allreduce(self, slot: int, t: torch.Tensor):
        """t (fp32, contiguous, <= SLOT_FLOATS elements) := sum over ranks of t, in place."""
        rc = _lib.lib().cris_peer_allreduce_f32(self.ptrs, self.world, self.rank, slot, t.data_ptr(), t.data_ptr(),
                                                t.numel(), self.timeout_s, _lib.stream_ptr())
        if rc != 0:
            raise RuntimeError(f"libcris_b200 cris_peer_allreduce_f32 failed: {_lib.lib().cris_last_error().decode()}")

    def bn_sync_fwd(self, site, partials, n_tiles, C, count, gamma, beta, eps, momentum, rm, rv, coef):
        """One fused exchange site (csrc/peer.cu peer_bn_sync_kernel): coef = [scale | shift | mean | invstd] (4C)."""
        p0 = coef.data_ptr()
        rc = _lib.lib().cris_peer_bn_sync_fwd(self.ptrs, self.world, self.rank, site, partials.data_ptr(), n_tiles, C,
                                              float(count), gamma.data_ptr(), beta.data_ptr(), eps, momentum,
                                              rm.data_ptr(), rv.data_ptr(), p0, p0 + 4 * C, p0 + 8 * C, p0 + 12 * C,
                                              self.timeout_s, _lib.stream_ptr())
        if rc != 0:
            raise RuntimeError(f"libcris_b200 cris_peer_bn_sync_fwd failed: {_lib.lib().cris_last_error().decode()}")

    def bn_sync_bwd(self, site, partials, n_tiles, C, sums_out, grad_beta, grad_gamma):
        rc = _lib.lib().cris_peer_bn_sync_bwd(self.ptrs, self.world, self.rank, site, partials.data_ptr(), n_tiles, C,
                                              sums_out.data_ptr(), grad_beta.data_ptr(), grad_gamma.data_ptr(),
                                              self.timeout_s, _lib.stream_ptr())
        if rc != 0:
            raise RuntimeError(f"libcris_b200 cris_peer_bn_sync_bwd failed: {_lib.lib().cris_last_error().decode()}")

    def close(self):
        L = _lib.lib()
        for r in range(self.world):
            p = self.ptrs[r]
            if p and r != self.rank:
                L.cris_peer_buffer_close(p, 0)
            self.ptrs[r] = None
        if self._own:
            L.cris_peer_buffer_close(self._own, 1)
            self._own = None


_exchange: dict = {}


def get_exchange(device: torch.device) -> Optional[PeerExchange]:
    """The process-wide exchange for `device` (created collectively on first use), or None when unavailable."""
    if os.environ.get("CRIS_B200_PEER", "1") == "0":
        return None
    key = device.index
    if key not in _exchange:
        ex = PeerExchange(device)
        if not ex.ok and dist.get_rank() == 0:
            print(f"[cris.pytorch_b200] NVLink peer exchange unavailable ({ex.error}); using torch.distributed",
                  flush=True)
        _exchange[key] = ex if ex.ok else None
    return _exchange[key]
