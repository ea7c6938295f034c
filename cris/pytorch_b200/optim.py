"""Adam with the whole parameter-group update in one kernel launch (csrc/optim.cu; SURVEY §8f "next" row).

Drop-in for the optimizer the reference builds at train.py:105-107 (`torch.optim.Adam(param_list, lr, weight_decay)`)
and drives through `GradScaler` (engine/engine.py:52-57): same constructor arguments, same `state_dict()` layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter, so checkpoints move between the two), same update arithmetic
(L2 weight decay folded into the gradient, bias-corrected moments), `MultiStepLR` works on `param_groups[i]["lr"]`.
`_step_supports_amp_scaling` makes `scaler.step(optimizer)` hand over the loss scale and the found-inf flag: the
kernel un-scales gradients on the fly, and a step with non-finite gradients is skipped entirely.
No CPU fallback: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, *,
                 maximize=False, decoupled_weight_decay=False):
        if amsgrad or maximize or decoupled_weight_decay:
            raise NotImplementedError("cris.pytorch_b200.optim.Adam: amsgrad / maximize / decoupled weight decay are "
                                      "not implemented (the reference uses none of them, train.py:105-107)")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        # the extra keys keep state_dict()s interchangeable with torch.optim.Adam
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._tables: Dict[int, dict] = {}
        self._fast: Dict[int, dict] = {}  # per group: armed fast path (cached table, common step count)
        self.table_refreshes = 0
        # GradScaler's found-inf flag of the previous step() and the parameters whose counters it advanced: the flag
        # is read on the host one step late, so step() never waits for the GPU (see _resolve_pending)
        self._pending = None
        self._flag_host = None
        self._flag_turn = 0

    def _table(self, gi: int, params: List[torch.Tensor], grads, ms, vs):
        """Device table of one group, refreshed only when a pointer changed (the caching allocator usually hands the
        freshly allocated .grad tensors the same blocks every step).  The refresh is an asynchronous copy from a
        small ring of pinned staging buffers, so it never makes the host wait for the backward pass in flight."""
        key = tuple(t.data_ptr() for ts in (params, grads, ms, vs) for t in ts)
        ent = self._tables.get(gi)
        if ent is not None and ent["key"] == key:
            return ent["dev"], ent["n"], ent["total"]
        L = _lib.lib()
        chunk = L.cris_adam_chunk_elems()
        n = len(params)
        if ent is None or ent["n"] != n:
            ent = {"dev": torch.empty((n, 6), dtype=torch.int64, device=params[0].device),
                   "pinned": [torch.empty((n, 6), dtype=torch.int64).pin_memory() for _ in range(4)],
                   "events": [None] * 4, "turn": 0, "n": n}
            assert L.cris_adam_table_entry_bytes() == 6 * 8
            self._tables[gi] = ent
        slot = ent["turn"] % 4
        ent["turn"] += 1
        if ent["events"][slot] is not None:
            ent["events"][slot].synchronize()  # that staging buffer's previous upload (4 refreshes ago) is done
        tab = ent["pinned"][slot].numpy()
        tab[:, 0] = [p.data_ptr() for p in params]
        tab[:, 1] = [g.data_ptr() for g in grads]
        tab[:, 2] = [m.data_ptr() for m in ms]
        tab[:, 3] = [v.data_ptr() for v in vs]
        tab[:, 4] = [p.numel() for p in params]
        chunks = (tab[:, 4] + chunk - 1) // chunk
        tab[:, 5] = np.cumsum(chunks) - chunks
        with torch.cuda.device(params[0].device):
            ent["dev"].copy_(ent["pinned"][slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        ent["events"][slot] = ev
        ent["key"], ent["total"] = key, int(chunks.sum())
        self.table_refreshes += 1
        return ent["dev"], n, ent["total"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        # GradScaler semantics: a step that saw inf/nan gradients changes nothing, not even the step counters.  The
        # kernel tests the device flag itself; the host-side counters are advanced optimistically and rolled back
        # when the flag is read at the next call (by then the GPU is long past it), so no step waits on the device.
        self._resolve_pending()
        advanced = []
        L = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            fast = self._fast.get(gi)
            params = group["params"]
            # Fast path (every step after the first): the same parameters, all with gradients at the same addresses
            # and one common step count.  The per-parameter `step` tensors of torch.optim.Adam's state layout are
            # brought up to date lazily (state_dict(), slow path): advancing 449 CPU tensors one by one cost ~1.6 ms
            # of host time per step, which DistributedDataParallel(find_unused_parameters=True) exposes on the GPU.
            if (fast is not None and fast["n"] == len(params) and all(p.grad is not None for p in params)
                    and params[0].grad.data_ptr() == fast["g0"] and params[-1].grad.data_ptr() == fast["g1"]
                    and params[len(params) // 2].grad.data_ptr() == fast["gm"]):
                fast["step"] += 1
                fast["lag"] += 1
                advanced.append(fast)
                tab, n, total = fast["table"]
                with torch.cuda.device(params[0].device):
                    rc = L.cris_adam_step(tab.data_ptr(), n, total, float(group["lr"]), float(group["betas"][0]),
                                          float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                          float(fast["step"]), grad_scale.data_ptr() if grad_scale is not None else None,
                                          found_inf.data_ptr() if found_inf is not None else None, _lib.stream_ptr())
                if rc != 0:
                    raise RuntimeError(f"libcris_b200 cris_adam_step failed: {L.cris_last_error().decode()}")
                torch._C._increment_version(params)
                continue
            self._sync_steps(gi)
            self._fast.pop(gi, None)
            by_step: Dict[float, list] = {}
            for p in params:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("cris.pytorch_b200.optim.Adam needs dense fp32 CUDA parameters and gradients "
                                       "(no CPU fallback)")
                st = self._state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                advanced.append(st)
                by_step.setdefault(float(st["step"]), []).append(p)
            for si, (step, plist) in enumerate(sorted(by_step.items())):
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in plist]
                if any(not p.is_contiguous() for p in plist):
                    raise RuntimeError("cris.pytorch_b200.optim.Adam needs contiguous parameters")
                ms = [self._state[p]["exp_avg"] for p in plist]
                vs = [self._state[p]["exp_avg_sq"] for p in plist]
                tab, n, total = self._table(gi * 1024 + si, plist, grads, ms, vs)
                with torch.cuda.device(plist[0].device):
                    rc = L.cris_adam_step(tab.data_ptr(), n, total, float(group["lr"]), float(group["betas"][0]),
                                          float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                          float(step), grad_scale.data_ptr() if grad_scale is not None else None,
                                          found_inf.data_ptr() if found_inf is not None else None, _lib.stream_ptr())
                if rc != 0:
                    raise RuntimeError(f"libcris_b200 cris_adam_step failed: {L.cris_last_error().decode()}")
                # the kernel wrote the parameters through raw pointers: tell autograd / every version-keyed cache
                # (engine.PackedWeights keeps bf16 copies keyed on p._version) that they changed in place
                torch._C._increment_version(plist)
            # arm the fast path when the whole group moved in lock step with contiguous gradients
            if len(by_step) == 1 and len(next(iter(by_step.values()))) == len(params) and \
                    all(p.grad.is_contiguous() for p in params):
                step0 = next(iter(by_step))
                ent = self._tables[gi * 1024]
                self._fast[gi] = {"n": len(params), "g0": params[0].grad.data_ptr(), "g1": params[-1].grad.data_ptr(),
                                  "gm": params[len(params) // 2].grad.data_ptr(), "step": step0, "lag": 0,
                                  "table": (ent["dev"], ent["n"], ent["total"])}
        if found_inf is not None:
            # the flag travels to pinned host memory behind this step's kernels; reading it next time waits on that
            # copy's event only (a plain .item() would drain the whole stream, i.e. the next backward pass)
            if self._flag_host is None:
                self._flag_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            host = self._flag_host[self._flag_turn & 1]
            self._flag_turn += 1
            with torch.cuda.device(found_inf.device):
                host.copy_(found_inf.reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self._pending = (host, ev, advanced)
        return loss

    def _sync_steps(self, gi=None):
        """Bring the per-parameter `step` tensors up to date with the fast path's group counters."""
        for g, fast in list(getattr(self, "_fast", {}).items()):
            if gi is not None and g != gi:
                continue
            if fast["lag"]:
                for p in self.param_groups[g]["params"]:
                    self._state[p]["step"] += fast["lag"]
                fast["lag"] = 0

    # `optimizer.state[p]["step"]` stays observable exactly like torch.optim.Adam's: any outside access to `.state`
    # first folds the fast path's pending counts into the per-parameter tensors (step() itself uses `_state`)
    @property
    def state(self):
        self._sync_steps()
        return self._state

    @state.setter
    def state(self, value):
        self._state = value

    def __setstate__(self, state):
        # torch's Optimizer.__setstate__ / load_state_dict write `state` straight into __dict__, past the property
        super().__setstate__(state)
        if "state" in self.__dict__:
            self._state = self.__dict__.pop("state")

    def _resolve_pending(self):
        if self._pending is not None:
            host, ev, advanced = self._pending
            self._pending = None
            ev.synchronize()
            if float(host[0]) != 0.0:
                for st in advanced:
                    if "lag" in st:      # a fast-path group record: roll its counters back
                        st["step"] -= 1
                        st["lag"] -= 1
                    else:
                        st["step"] -= 1

    def state_dict(self):
        self._resolve_pending()
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        self._resolve_pending()
        self._fast.clear()
        return super().load_state_dict(state_dict)
