"""The sm_100a execution engine behind `cris.pytorch_b200.CRIS`.

One forward of the reference's `CRIS.forward` (model/segmenter.py:29-62) is a fixed sequence of launches of
the hand-written kernels in libcris_b200.so over torch-allocated device buffers; the backward replays a tape of
closures recorded during the forward (a purpose-built reverse-mode pass: PyTorch's autograd only sees one
`torch.autograd.Function`, so DDP / GradScaler / optimizers of the reference's train.py keep working).

Layout: image activations are bf16 "padded NHWC" row matrices [N*(H+2)*(W+2), C] with a zero border, so that a
3x3 convolution is 9 row-shifted GEMM taps; token activations are [B*L, C] (bf16, or fp32 for
the transformer residual streams).  PyTorch is used for memory, streams and torch.distributed only.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import GemmArgs, call, gemm

ACT_NONE, ACT_RELU, ACT_QGELU, ACT_RELU_POST = 0, 1, 2, 3
TAP_NONE, TAP_ACCUM, TAP_WGRAD = 0, 1, 2
BN_EPS, LN_EPS, BN_MOMENTUM = 1e-5, 1e-5, 0.1
PEER_SLOT_FLOATS, PEER_MAX_SLOTS = 4096, 1024  # CRIS_PEER_* in include/cris_b200.h


def _r8(x: int) -> int:
    return (x + 7) // 8 * 8


class Mat:
    """A [rows, C] device matrix view (pitch `ld` elements) over a torch buffer; optional padded-NHWC geometry."""
    __slots__ = ("buf", "ptr", "rows", "C", "ld", "fp32", "geom", "root", "col0", "gbuf", "gwritten", "need_grad")

    def __init__(self, buf, rows, C, ld=None, fp32=False, geom=None, ptr=None, root=None, col0=0):
        self.buf, self.rows, self.C = buf, rows, C
        self.ld = C if ld is None else ld
        self.fp32, self.geom = fp32, geom
        self.ptr = buf.data_ptr() if ptr is None else ptr
        self.root, self.col0 = root, col0
        self.gbuf: Optional["Mat"] = None
        self.gwritten = False
        self.need_grad = True

    @property
    def esize(self):
        return 4 if self.fp32 else 2

    def cols(self, c0: int, c1: int) -> "Mat":
        r = self.root if self.root is not None else self
        return Mat(self.buf, self.rows, c1 - c0, self.ld, self.fp32, self.geom, self.ptr + c0 * self.esize, r,
                   self.col0 + c0)

    def rows_slice(self, r0: int, r1: int) -> "Mat":
        m = Mat(self.buf, r1 - r0, self.C, self.ld, self.fp32, None, self.ptr + r0 * self.ld * self.esize)
        return m

    @property
    def hp(self):
        return self.geom[1] + 2 if self.geom else 0

    @property
    def wp(self):
        return self.geom[2] + 2 if self.geom else 0


class PackedWeights:
    """bf16 kernel-layout copies of the fp32 master parameters, refreshed when a parameter's version changes
    (optimizer steps are in-place): conv 3x3 -> [Cout][9][cin_pad]; 1x1 conv / Linear / in_proj -> [out][in]."""

    def __init__(self):
        self.cache: Dict[str, tuple] = {}
        self.force = False        # CUDA-graph capture: always (re)launch the pack kernel into the cached buffer
        self.done_in_pass = set()  # ... but only once per captured pass

    def _multi_entries(self):
        """[(cache key, Mat, parameter, scale pointer or 0)]: plain copies, and (eval) BatchNorm-folded copies"""
        out = []
        for k, v in self.cache.items():
            if k.endswith("#folded"):
                if getattr(self, "_multi_folded", False):
                    out.append((k, v[0], v[1], v[2]))
            elif len(v) == 3:
                out.append((k, v[1], v[2], 0))
        return out

    def _multi_sig(self, ents):
        return tuple((p.data_ptr(), m.buf.data_ptr(), sc) for _, m, p, sc in ents)

    def prepare_multi(self, include_folded: bool = False):
        """Build / refresh the device table of refresh_all() — call OUTSIDE a CUDA-graph capture (it copies host data).
        include_folded: the eval graph also refreshes the BatchNorm-folded copies."""
        self._multi_folded = include_folded
        ents = self._multi_entries()
        if not ents:
            self._multi_key = None
            return
        key = self._multi_sig(ents)
        if getattr(self, "_multi_key", None) == key:
            return
        import numpy as np
        L = _lib.lib()
        chunk = L.cris_pack_chunk_elems()
        assert L.cris_pack_entry_bytes() == 56
        tab = np.zeros((len(ents), 7), dtype=np.int64)
        tot = 0
        for i, (_, m, p, sc) in enumerate(ents):
            w = p.detach()
            if w.dim() == 4 and w.shape[2] == 3 and m.C == 9 * _r8(w.shape[1]):
                rows, cols, ld, taps = w.shape[0], w.shape[1], _r8(w.shape[1]), 9
            else:
                rows, cols, ld, taps = w.shape[0], w.numel() // w.shape[0], m.ld, 1
            tab[i, 0], tab[i, 1], tab[i, 2] = w.data_ptr(), m.buf.data_ptr(), rows
            tab[i, 3] = cols | (ld << 32)            # int32 cols, int32 ld
            tab[i, 4] = taps                          # int32 taps, int32 pad
            tab[i, 5] = tot
            tab[i, 6] = sc
            tot += (rows * taps * ld + chunk - 1) // chunk
        self._multi_tab = torch.from_numpy(tab).to(ents[0][1].buf.device)
        self._multi_key, self._multi_total, self._multi_n = key, tot, len(ents)

    def refresh_all(self) -> bool:
        """Re-pack EVERY cached copy with one kernel launch (csrc/tokens.cu pack_multi_kernel) and mark them done for
        this pass — used inside CUDA-graph captures, where every copy must be refreshed on every replay (one launch
        instead of ~143).  Needs prepare_multi() to have run on the current set of copies; otherwise the per-tensor
        launches of get() take over."""
        ents = self._multi_entries()
        key = self._multi_sig(ents) if ents else None
        if not ents or getattr(self, "_multi_key", None) != key:
            return False
        call("cris_pack_multi", self._multi_tab.data_ptr(), self._multi_n, self._multi_total)
        for e in ents:
            self.done_in_pass.add(e[0])
        return True

    def get(self, name: str, p: torch.Tensor, as_matrix: bool = False) -> Mat:
        ent = self.cache.get(name)
        key = (p._version, p.data_ptr())
        if ent is not None and ent[0] == key and not (self.force and name not in self.done_in_pass):
            return ent[1]
        self.done_in_pass.add(name)
        w = p.detach()
        if w.dim() == 4 and w.shape[2] == 3 and not as_matrix:
            cout, cin = w.shape[0], w.shape[1]
            cp = _r8(cin)
            buf = ent[1].buf if ent is not None else torch.empty(cout, 9 * cp, dtype=torch.bfloat16, device=w.device)
            call("cris_pack_conv_weight", w.data_ptr(), buf.data_ptr(), cout, cin, 9, cp)
            m = Mat(buf, cout, 9 * cp)
        else:
            rows = w.shape[0]
            cols = w.numel() // rows
            ld = _r8(cols)
            buf = ent[1].buf if ent is not None else torch.empty(rows, ld, dtype=torch.bfloat16, device=w.device)
            call("cris_pack_matrix", w.data_ptr(), buf.data_ptr(), rows, cols, ld)
            m = Mat(buf, rows, cols, ld)
        self.cache[name] = (key, m, p)
        return m

    def get_folded(self, name: str, p: torch.Tensor, scale_ptr: int, as_matrix: bool = False) -> Mat:
        """Eval mode: the same layouts with output channel co scaled by scale[co] (BatchNorm folded into the
        convolution).  Re-packed on every call (the scale depends on the running statistics, not on p._version) unless
        refresh_all() already did it for this captured pass."""
        ent = self.cache.get(name + "#folded")
        if ent is not None and ent[2] == scale_ptr and self.force and (name + "#folded") in self.done_in_pass:
            return ent[0]
        w = p.detach()
        if w.dim() == 4 and w.shape[2] == 3 and not as_matrix:
            cout, cin = w.shape[0], w.shape[1]
            cp = _r8(cin)
            buf = ent[0].buf if ent is not None else torch.empty(cout, 9 * cp, dtype=torch.bfloat16, device=w.device)
            call("cris_pack_conv_weight_scaled", w.data_ptr(), scale_ptr, buf.data_ptr(), cout, cin, 9, cp)
            m = Mat(buf, cout, 9 * cp)
        else:
            rows = w.shape[0]
            cols = w.numel() // rows
            ld = _r8(cols)
            buf = ent[0].buf if ent is not None else torch.empty(rows, ld, dtype=torch.bfloat16, device=w.device)
            call("cris_pack_matrix_scaled", w.data_ptr(), scale_ptr, buf.data_ptr(), rows, cols, ld)
            m = Mat(buf, rows, cols, ld)
        self.cache[name + "#folded"] = (m, p, scale_ptr)
        return m


def _bicubic_matrix(src: int, H: int, W: int) -> torch.Tensor:
    """[H*W, src*src] interpolation matrix of F.interpolate(mode='bicubic', align_corners=False) from a
    src x src grid (model/clip.py:101-104).  Batch-independent constant, built once on the host."""
    eye = torch.eye(src * src).reshape(src * src, 1, src, src)
    out = F.interpolate(eye, size=(H, W), mode="bicubic", align_corners=False)  # [s*s, 1, H, W]
    return out.reshape(src * src, H * W).t().contiguous()


def _pos1d(d: int, length: int) -> torch.Tensor:
    """model/layers.py:106-123 — [length, d] sinusoid table (constant)."""
    pos = torch.arange(length, dtype=torch.float32)[:, None]
    freq = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(length, d)
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe


def _pos2d(d: int, H: int, W: int) -> torch.Tensor:
    """model/layers.py:125-152 — [H*W, d]: first half of the channels encodes w, second half h (constant)."""
    half = d // 2
    freq = torch.exp(torch.arange(0.0, half, 2) * -(math.log(10000.0) / half))
    pw = torch.arange(0.0, W)[:, None] * freq
    ph = torch.arange(0.0, H)[:, None] * freq
    pe = torch.zeros(H, W, d)
    pe[:, :, 0:half:2] = torch.sin(pw)[None]
    pe[:, :, 1:half:2] = torch.cos(pw)[None]
    pe[:, :, half::2] = torch.sin(ph)[:, None]
    pe[:, :, half + 1::2] = torch.cos(ph)[:, None]
    return pe.reshape(H * W, d)


def mat_to_torch(m: Mat) -> torch.Tensor:
    """Debug/test helper: copy a Mat out as fp32 — NCHW for padded-NHWC image tensors, [rows, C] otherwise."""
    esz = m.esize
    flat = m.buf.reshape(-1)
    off = (m.ptr - m.buf.data_ptr()) // esz
    t = torch.as_strided(flat, (m.rows, m.C), (m.ld, 1), off).float()
    if m.geom is not None:
        N, H, W = m.geom
        t = t.reshape(N, H + 2, W + 2, m.C)[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).contiguous()
    return t


class Engine:
    def __init__(self, model: nn.Module):
        self.model = model
        self.packed = PackedWeights()
        self.consts: Dict[tuple, torch.Tensor] = {}
        self.step = 0
        self.debug_taps: Optional[dict] = None  # set to {} to collect named intermediates (tests only)
        # bench.py: CUDA-event pairs recorded around the forward GEMM launch of one named convolution, so the
        # dominant kernel is timed live inside the timed step (on the stream it is launched on)
        self.probe_name: Optional[str] = None
        self.probe_events: List = []
        # whole-pass CUDA graphs (forward graph + backward graph per input shape): removes ~2000 host launches/step
        self.use_graphs = os.environ.get("CRIS_B200_GRAPHS", "1") != "0"
        self.gemm_log: Optional[list] = None  # profiling: (M,N,K,batch,...) of every GEMM launch, in order
        self.force_sync_bn = False  # tests: exercise the cross-rank BN exchange without SyncBatchNorm modules
        # eval mode: BatchNorm folded into the producing convolution (weights x scale, shift as bias, residual + ReLU in
        # the GEMM epilogue) instead of a bn_coeffs + bn_apply pass per layer
        self.fold_eval_bn = os.environ.get("CRIS_B200_FOLD_EVAL_BN", "1") != "0"
        self.last_metric_counts: Optional[torch.Tensor] = None  # int32 [B,2] of the last training forward
        self.eval_coefs: Dict[str, torch.Tensor] = {}  # BN prefix -> persistent [scale|shift|mean|invstd] (eval fold)
        self._bn_multi = None
        self.graphs: Dict[tuple, "GraphedStep"] = {}
        self.eval_graphs: Dict[tuple, "GraphedEval"] = {}
        self._counter: Optional[torch.Tensor] = None
        _lib.lib()

    @classmethod
    def bare(cls) -> "Engine":
        """An engine without a model: used by the per-op tests to drive single building blocks."""
        class _NoModel(nn.Module):
            dropout_p = 0.0
        return cls(_NoModel())

    def needs_sync_bn(self) -> bool:
        """Training statistics are shared across ranks iff the model was converted to SyncBatchNorm (train.py:97-98)."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return False
        return self.force_sync_bn or any(isinstance(m, nn.SyncBatchNorm) for m in self.model.modules())

    def peer_exchange(self, device):
        """NVLink peer-memory exchange (collective on first call), or None -> torch.distributed.all_reduce."""
        if device.type != "cuda":
            return None
        from . import peer
        return peer.get_exchange(device)

    def eval_coef(self, prefix: str, C: int, device) -> torch.Tensor:
        t = self.eval_coefs.get(prefix)
        if t is None or t.device != device or t.numel() != 4 * C:
            t = torch.empty(4 * C, dtype=torch.float32, device=device)
            self.eval_coefs[prefix] = t
        return t

    def prepare_bn_multi(self):
        """Device table for cris_bn_coeffs_multi over every BatchNorm that eval_coef() has seen (outside captures)."""
        if not self.eval_coefs:
            self._bn_multi = None
            return
        import numpy as np
        P, Bf = dict(self.model.named_parameters()), dict(self.model.named_buffers())
        names = sorted(self.eval_coefs)
        key = tuple((self.eval_coefs[n].data_ptr(), P[n + ".weight"].data_ptr(), Bf[n + ".running_mean"].data_ptr()) for n in names)
        if self._bn_multi is not None and self._bn_multi[0] == key:
            return
        assert _lib.lib().cris_bn_eval_entry_bytes() == 48
        tab = np.zeros((len(names), 6), dtype=np.int64)
        blocks = 0
        for i, n in enumerate(names):
            C = P[n + ".weight"].numel()
            tab[i, 0], tab[i, 1] = P[n + ".weight"].data_ptr(), P[n + ".bias"].data_ptr()
            tab[i, 2], tab[i, 3] = Bf[n + ".running_mean"].data_ptr(), Bf[n + ".running_var"].data_ptr()
            tab[i, 4] = self.eval_coefs[n].data_ptr()
            tab[i, 5] = C | (blocks << 32)
            blocks += (C + 127) // 128
        dev = self.eval_coefs[names[0]].device
        self._bn_multi = (key, torch.from_numpy(tab).to(dev), len(names), blocks, set(names))

    def side_stream(self, device) -> "torch.cuda.Stream":
        st = getattr(self, "_side", None)
        if st is None or st.device != device:
            st = torch.cuda.Stream(device=device)
            self._side = st
        return st

    def wgrad_stream(self, device) -> "torch.cuda.Stream":
        """Lowest-priority stream for the weight-gradient branch of the captured backward (Run.leaf_branch)."""
        st = getattr(self, "_wgrad_side", None)
        if st is None or st.device != device:
            st = torch.cuda.Stream(device=device, priority=0)
            self._wgrad_side = st
        return st

    def step_counter(self, device) -> torch.Tensor:
        if self._counter is None or self._counter.device != device:
            self._counter = torch.zeros(1, dtype=torch.int64, device=device)
        return self._counter

    def const(self, key, fn, device):
        t = self.consts.get(key)
        if t is None or t.device != device:
            t = fn().to(device)
            self.consts[key] = t
        return t

    # -------------------------------------------------------------------------------------------
    def run(self, img, word, mask):
        model = self.model
        training = model.training
        if training and mask is None:
            raise ValueError("CRIS.forward in training mode needs the mask (model/segmenter.py:54-59)")
        _lib.device_check()
        if training and torch.is_grad_enabled():
            names, params = [], []
            for k, p in model.named_parameters():
                if k == "backbone.logit_scale":  # never used by the path
                    continue
                names.append(k)
                params.append(p)
            # cross-rank BatchNorm statistics inside a captured graph need the NVLink peer exchange (csrc/peer.cu):
            # a NCCL call cannot be replayed from the forward/backward graphs
            graphs_ok = self.use_graphs and (not self.needs_sync_bn() or self.peer_exchange(img.device) is not None
                                             or os.environ.get("CRIS_B200_GRAPHS_DDP", "0") == "1")
            if graphs_ok and self.debug_taps is None and self.probe_name is None and _lib._prof is None:
                key = (tuple(img.shape), tuple(word.shape), tuple(mask.shape), img.device.index, float(model.dropout_p),
                       self._param_fingerprint())
                gs = self.graphs.get(key)
                if gs is None:
                    gs = GraphedStep(self, img, word, mask, names)
                    self.graphs = {key: gs}  # one shape resident at a time (each holds all activations)
                if len(gs.gbs) > 1:
                    P = dict(zip(names, params))
                    K = len(gs.gbs)
                    pred, mask_r, loss = _GraphFirst.apply(gs, img, word, mask, *[P[n] for n in gs.seg_names[K - 1]])
                    for k in range(K - 2, -1, -1):
                        loss = _GraphSeg.apply(gs, k, loss, *[P[n] for n in gs.seg_names[k]])
                    return pred, mask_r, loss
                return _GraphFunction.apply(gs, img, word, mask, *params)
            pred, mask_r, loss = _CRISFunction.apply(self, names, img, word, mask, *params)
            return pred, mask_r, loss
        if (not training and self.use_graphs and self.debug_taps is None and self.probe_name is None
                and self.gemm_log is None and _lib._prof is None and img.is_cuda):
            key = (tuple(img.shape), tuple(word.shape), img.device.index, self._param_fingerprint())
            ge = self.eval_graphs.get(key)
            if ge is None:
                ge = GraphedEval(self, img, word)
                self.eval_graphs = {key: ge}
            return ge(img, word)
        with torch.no_grad():
            r = Run(self, img, word, mask, training, record=False)
            r.forward()
        if training:
            return r.pred, r.mask_out, r.loss
        return r.pred

    def _param_fingerprint(self):
        ps = list(self.model.parameters())
        return (len(ps), ps[0].data_ptr(), ps[-1].data_ptr()) if ps else (0, 0, 0)


class _CRISFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, names, img, word, mask, *params):
        r = Run(engine, img, word, mask, True, record=True)
        r.forward()
        ctx.run = r
        ctx.names = names
        ctx.mark_non_differentiable(r.pred, r.mask_out)
        return r.pred, r.mask_out, r.loss

    @staticmethod
    def backward(ctx, _dpred, _dmask, dloss):
        r: Run = ctx.run
        grads = r.backward(dloss, ctx.names)
        ctx.run = None
        return (None, None, None, None, None, *grads)


class GraphedStep:
    """One training pass (forward graph + backward graph) captured for a fixed input shape.

    All activations, gradients and workspaces live in the graphs' private memory pool at fixed addresses; inputs
    are copied into static buffers, the loss gradient into a static scalar, and both graphs are replayed.
    Dropout masks stay fresh because kernels add a device-side step counter to their seeds."""

    def __init__(self, engine: Engine, img, word, mask, names):
        self.names = names
        dev = img.device
        self.img = img.detach().float().contiguous().clone()
        self.word = word.detach().long().contiguous().clone()
        self.mask = mask.detach().float().contiguous().clone()
        self.g = torch.ones(1, dtype=torch.float32, device=dev)
        # eager warm-up on a side stream (fills constant caches, sets kernel attributes); BatchNorm buffers are
        # restored afterwards so that the warm-up does not count as a training step
        saved = {k: b.clone() for k, b in engine.model.named_buffers()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            r = Run(engine, self.img, self.word, self.mask, True, record=True)
            r.forward()
            n_tape = len(r.tape)
            r.backward(self.g, names)
            touched = dict(r.touched)
            del r
        torch.cuda.current_stream(dev).wait_stream(side)
        # backward segments (CRIS_B200_BWD_SEGMENTS=K): the reversed tape is cut into K ranges captured as K graphs;
        # a parameter belongs to the range that writes its gradient last (learnt from the warm-up pass above)
        K = max(1, min(int(os.environ.get("CRIS_B200_BWD_SEGMENTS", "1")), 8, n_tape))
        self.bounds = [round(n_tape * k / K) for k in range(K + 1)]
        self.seg_names: List[List[str]] = [[] for _ in range(K)]
        for nm in names:
            pos = touched.get(nm, n_tape - 1)
            k = next(j for j in range(K) if pos < self.bounds[j + 1] or j == K - 1)
            self.seg_names[k].append(nm)
        if any(len(sn) == 0 for sn in self.seg_names):  # every link of the autograd chain needs a parameter
            self.bounds, self.seg_names = [0, n_tape], [list(names)]
        with torch.no_grad():
            for k, b in engine.model.named_buffers():
                b.copy_(saved[k])
        torch.cuda.synchronize(dev)
        engine.packed.prepare_multi()  # device table of the one-launch weight repack (host copy: outside the capture)
        engine.packed.force = True
        try:
            with torch.no_grad():
                self.gf = torch.cuda.CUDAGraph()
                engine.packed.done_in_pass = set()
                n0 = _lib.launch_count()
                with torch.cuda.graph(self.gf):
                    r = Run(engine, self.img, self.word, self.mask, True, record=True)
                    r.forward()
                n1 = _lib.launch_count()
                self.pred, self.mask_out, self.loss = r.pred, r.mask_out, r.loss
                self.metric_counts = r.metric_counts
                self.engine = engine
                self.gbs, self.n_bwds = [], []
                for k in range(len(self.seg_names)):
                    gb = torch.cuda.CUDAGraph()
                    nb0 = _lib.launch_count()
                    # the dependent chain is captured on a high-priority stream: where the weight-gradient branch
                    # (Run.leaf_branch, default-priority stream) competes for SMs, the chain's thread blocks go first
                    cap = (torch.cuda.Stream(device=dev, priority=-1)
                           if (os.environ.get("CRIS_B200_WGRAD_STREAM", "1") == "1"
                               and os.environ.get("CRIS_B200_WGRAD_PRIO", "1") == "1") else None)
                    with torch.cuda.graph(gb, pool=self.gf.pool(), **({"stream": cap} if cap is not None else {})):
                        r.backward_range(self.g, self.bounds[k], self.bounds[k + 1] if k + 1 < len(self.seg_names) else None)
                    self.gbs.append(gb)
                    self.n_bwds.append(_lib.launch_count() - nb0)
                self.grads = r.collect_grads(names)
                self.grad_of_name = dict(zip(names, self.grads))
                P_ = dict(engine.model.named_parameters())
                self.params = [P_[n] for n in names]
                self.param_of_name = dict(zip(names, self.params))
                self.gb = self.gbs[0]
                self.n_fwd, self.n_bwd = n1 - n0, _lib.launch_count() - n1  # kernels inside each graph
                _lib.lib().cris_add_launch_count(-(self.n_fwd + self.n_bwd) & ((1 << 64) - 1))  # capture != launch
                self.run = r  # keeps every captured buffer referenced
                self._tables = getattr(engine.packed, "_multi_tab", None)  # the captured pack launch reads this table
            import gc
            gc.collect()
            if os.environ.get("CRIS_B200_GC_FREEZE", "1") == "1":  # process-wide side effect: opt-out documented in INTEGRATION.md
                gc.freeze()  # the captured pass holds ~10^5 long-lived objects: keep them out of later GC passes
        finally:
            engine.packed.force = False


def _fresh_grads(self, names=None):
    """Gradients handed to autograd.  AccumulateGrad STEALS a gradient tensor only when nobody else references the
    tensor object; handing it the static graph buffers themselves (referenced from this object) made it clone every
    one of the 449 gradients on every step — 2.1 ms of copy kernels after the backward graph (tools/bwd_overhead.py).
    A fresh alias (`detach()`: new tensor object, same storage) is stolen instead, so `p.grad` IS the graph's buffer.
    That is only safe while nothing accumulates into an existing `.grad` (the next replay overwrites the buffer), so
    if any parameter still holds a gradient — gradient accumulation without zero_grad(set_to_none=True) — the
    copying path is taken."""
    gl = self.grads if names is None else [self.grad_of_name[n] for n in names]
    ps = self.params if names is None else [self.param_of_name[n] for n in names]
    if any(p.grad is not None for p in ps):
        return [g.clone() for g in gl]
    return [g.detach() for g in gl]


GraphedStep.fresh_grads = _fresh_grads


class GraphedEval:
    """The inference pass (model.eval(): engine/engine.py:100,171, tools/latency.py:62) captured for one input
    shape: ~700 kernel launches become one graph launch, which is what bounds batch-1 latency.  Parameters and
    BatchNorm running statistics are read at replay time (weight re-packing is part of the graph)."""

    def __init__(self, engine: "Engine", img, word):
        dev = img.device
        self.img = img.detach().float().contiguous().clone()
        self.word = word.detach().long().contiguous().clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            r = Run(engine, self.img, self.word, None, False, record=False)
            r.forward()
            del r
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        engine.prepare_bn_multi()
        engine.packed.prepare_multi(include_folded=True)
        engine.packed.force = True
        try:
            with torch.no_grad():
                self.g = torch.cuda.CUDAGraph()
                engine.packed.done_in_pass = set()
                n0 = _lib.launch_count()
                with torch.cuda.graph(self.g):
                    r = Run(engine, self.img, self.word, None, False, record=False)
                    r.forward()
                self.n = _lib.launch_count() - n0
                _lib.lib().cris_add_launch_count(-self.n & ((1 << 64) - 1))  # capture != launch
                self.pred = r.pred
                # device tables the captured multi-tensor launches read: keep them alive as long as the graph
                self._tables = (getattr(engine.packed, "_multi_tab", None), engine._bn_multi)
        finally:
            engine.packed.force = False

    def __call__(self, img, word):
        self.img.copy_(img)
        self.word.copy_(word)
        self.g.replay()
        _lib.lib().cris_add_launch_count(self.n)
        return self.pred.clone()  # callers may keep predictions across iterations (engine.py:171-190)


class _GraphFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gs: GraphedStep, img, word, mask, *params):
        prof = os.environ.get("CRIS_B200_HOSTPROF") == "1"
        if prof:
            import time
            t0 = time.perf_counter()
        gs.img.copy_(img)
        gs.word.copy_(word)
        gs.mask.copy_(mask)
        if prof:
            t1 = time.perf_counter()
        gs.gf.replay()
        if prof:
            t2 = time.perf_counter()
            if t2 - t0 > 0.01:
                print(f"[hostprof] input copies {1e3 * (t1 - t0):.1f} ms, graph launch {1e3 * (t2 - t1):.1f} ms", flush=True)
        _lib.lib().cris_add_launch_count(gs.n_fwd)
        gs.engine.last_metric_counts = gs.metric_counts  # refilled by this replay (read by CRIS.train_metric())
        ctx.gs = gs
        # fresh tensors: the static capture buffers are overwritten by the next replay, and callers may keep
        # predictions / masks across iterations (metric accumulation), as they can with the reference
        pred, mask_out, loss = gs.pred.detach().clone(), gs.mask_out.detach().clone(), gs.loss.detach().clone()
        ctx.mark_non_differentiable(pred, mask_out)
        return pred, mask_out, loss

    @staticmethod
    def backward(ctx, _dpred, _dmask, dloss):
        gs: GraphedStep = ctx.gs
        gs.g.copy_(dloss.detach().float().reshape(1))
        gs.gb.replay()
        _lib.lib().cris_add_launch_count(gs.n_bwd)
        return (None, None, None, None, *gs.fresh_grads())


class _GraphFirst(torch.autograd.Function):
    """Segmented backward, first link of the chain: replays the forward graph; its backward replays the LAST
    backward range (the earliest layers) and returns the gradients finalised there."""

    @staticmethod
    def forward(ctx, gs: GraphedStep, img, word, mask, *params):
        gs.img.copy_(img)
        gs.word.copy_(word)
        gs.mask.copy_(mask)
        gs.gf.replay()
        _lib.lib().cris_add_launch_count(gs.n_fwd)
        gs.engine.last_metric_counts = gs.metric_counts
        ctx.gs = gs
        pred, mask_out, loss = gs.pred.detach().clone(), gs.mask_out.detach().clone(), gs.loss.detach().clone()
        ctx.mark_non_differentiable(pred, mask_out)
        return pred, mask_out, loss

    @staticmethod
    def backward(ctx, _dpred, _dmask, _dtoken):
        gs: GraphedStep = ctx.gs
        k = len(gs.gbs) - 1
        gs.gbs[k].replay()
        _lib.lib().cris_add_launch_count(gs.n_bwds[k])
        return (None, None, None, None, *gs.fresh_grads(gs.seg_names[k]))


class _GraphSeg(torch.autograd.Function):
    """Segmented backward, link k < K-1: identity on the loss token in the forward; the backward replays backward
    range k (range 0 = loss head + last layers) and hands its parameters' gradients to autograd — and therefore
    to DistributedDataParallel's bucket all-reduce — before the remaining ranges run."""

    @staticmethod
    def forward(ctx, gs: GraphedStep, k: int, token, *params):
        ctx.gs, ctx.k = gs, k
        return token.view_as(token)

    @staticmethod
    def backward(ctx, dtoken):
        gs: GraphedStep = ctx.gs
        k = ctx.k
        if k == 0:
            gs.g.copy_(dtoken.detach().float().reshape(1))
        gs.gbs[k].replay()
        _lib.lib().cris_add_launch_count(gs.n_bwds[k])
        return (None, None, dtoken, *gs.fresh_grads(gs.seg_names[k]))


class Run:
    """State of one forward (+ backward) pass."""
    # defaults for helpers that tests build without __init__
    xchg = None
    xslot = 0
    _zarena: Optional[torch.Tensor] = None
    _zoff = 0
    bwd_pos = 0
    touched: Optional[Dict[str, int]] = None
    in_backward = False

    def __init__(self, engine: Engine, img, word, mask, training: bool, record: bool):
        self.e = engine
        self.model = engine.model
        self.dev = img.device
        self.img = img.contiguous().float()
        self.word = word.contiguous().long()
        self.mask = None if mask is None else mask.contiguous().float()
        self.training = training
        self.record = record
        self.tape: List = []
        self.P = dict(self.model.named_parameters())
        self.Bf = dict(self.model.named_buffers())
        self.pgrad: Dict[str, torch.Tensor] = {}
        self.p_drop = float(self.model.dropout_p) if training else 0.0
        engine.step += 1
        # dropout masks = hash(site seed + device step counter, element index): the counter lives on the device and
        # is bumped by a kernel at the start of every forward, so a replayed CUDA graph draws fresh masks
        self.seed_base = (torch.initial_seed() * 1000003) & ((1 << 61) - 1)
        self.n_seed = 0
        self.seed_dev = engine.step_counter(self.dev).data_ptr()
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.sync_bn = training and engine.needs_sync_bn()
        self.xchg = engine.peer_exchange(self.dev) if self.sync_bn else None
        self.xslot = 0  # exchange site index inside this pass (forward sites, then backward sites)
        self._zarena: Optional[torch.Tensor] = None
        self._zoff = 0
        self.touched: Dict[str, int] = {}
        self.bwd_pos = 0

    # ---- small helpers -------------------------------------------------------------------------
    def new(self, rows, C, fp32=False, geom=None, zero=False, ld=None) -> Mat:
        ld = C if ld is None else ld
        dt = torch.float32 if fp32 else torch.bfloat16
        buf = (torch.zeros if zero else torch.empty)(rows, ld, dtype=dt, device=self.dev)
        return Mat(buf, rows, C, ld, fp32, geom)

    def padded(self, N, H, W, C, zero=False, ld=None) -> Mat:
        return self.new(N * (H + 2) * (W + 2), C, False, (N, H, W), zero, ld)

    def f32(self, n, zero=False):
        if zero:
            return self.zeros_f32(n)
        return torch.empty(n, dtype=torch.float32, device=self.dev)

    def zeros_f32(self, n):
        """n zero floats carved from a pre-zeroed arena (one memset per 4M floats instead of one per use).  One arena
        per CUDA stream: the memset is ordered only with the stream it was issued on (the text tower runs on a second
        stream inside graph captures)."""
        n_al = (n + 63) // 64 * 64
        sid = torch.cuda.current_stream(self.dev).cuda_stream if self.dev.type == "cuda" else 0
        arenas = self.__dict__.setdefault("_zarenas", {})
        ar = arenas.get(sid)
        if ar is None or ar[1] + n_al > ar[0].numel():
            ar = [torch.zeros(max(n_al, 1 << 22), dtype=torch.float32, device=self.dev), 0]
            arenas[sid] = ar
        t = ar[0][ar[1]:ar[1] + n]
        ar[1] += n_al
        return t

    def tap(self, name: str, m: Mat):
        if self.e.debug_taps is not None:
            self.e.debug_taps[name] = mat_to_torch(m)

    def seed(self):
        self.n_seed += 1
        return (self.seed_base + self.n_seed * 104729) & ((1 << 63) - 1)

    def w(self, name, as_matrix=False) -> Mat:
        return self.e.packed.get(name, self.P[name], as_matrix)

    def pg(self, name) -> torch.Tensor:
        """fp32 gradient buffer of a parameter: a view of ONE zero-filled flat buffer (one memset per backward)."""
        if self.touched is None:
            self.touched = {}
        self.touched[name] = self.bwd_pos  # last backward-tape position that writes this gradient
        g = self.pgrad.get(name)
        if g is None:
            if not self.pgrad:
                self.ensure_pgrad()
                g = self.pgrad.get(name)
            if g is None:
                g = torch.zeros_like(self.P[name], dtype=torch.float32)
                self.pgrad[name] = g
        return g

    def ensure_pgrad(self):
        """Allocate + zero the flat gradient buffer on the CURRENT stream (called before any branch forks)."""
        if self.pgrad:
            return
        total, offs = 0, {}
        for k, p in self.P.items():
            offs[k] = total
            total += (p.numel() + 3) // 4 * 4  # keep every view 16-byte aligned
        flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        for k, p in self.P.items():
            self.pgrad[k] = flat[offs[k]:offs[k] + p.numel()].view(p.shape)

    def leaf_branch(self, fn, hold=()):
        """Run a LEAF of the backward (a weight / bias gradient: nothing downstream in this pass reads its result) as a
        parallel branch of the captured graph: issued on a second, lower-priority stream forked from the current one, so
        its persistent kernels fill the launch gaps and tails of the dependent chain (dgrad -> BatchNorm backward ->
        dgrad ...) instead of sitting in it.  `hold` are the tensors the leaf reads: they stay referenced until the
        branch is joined (join_leaves), because the chain could otherwise free and re-use their memory while the leaf
        still runs.  Eager (non-captured) execution and the text-tower branch run the leaf in line."""
        st = getattr(self, "_leaf_main", None)
        if (st is None or self.dev.type != "cuda" or not torch.cuda.is_current_stream_capturing()
                or torch.cuda.current_stream(self.dev) != st):
            fn()
            return
        ws = self.e.wgrad_stream(self.dev)
        ws.wait_stream(st)
        with torch.cuda.stream(ws):
            fn()
        self._leaf_hold.extend(hold)
        self._leaf_pending += 1
        if self._leaf_pending >= self._leaf_join_every:
            self.join_leaves()

    def join_leaves(self):
        if getattr(self, "_leaf_pending", 0):
            self._leaf_main.wait_stream(self.e.wgrad_stream(self.dev))
            self._leaf_pending = 0
        if getattr(self, "_leaf_hold", None):
            self._leaf_hold.clear()

    def on_backward(self, fn):
        if self.record:
            self.tape.append(fn)

    # gradient slots ---------------------------------------------------------------------------
    def grad_slot(self, m: Mat):
        """-> (grad Mat, accumulate?) for the tensor `m`; handles column slices of concat buffers."""
        root = m.root if m.root is not None else m
        if root.gbuf is None:
            whole = m.root is None
            g = self.new(root.rows, root.C, root.fp32, root.geom, zero=not whole, ld=root.ld)
            root.gbuf = g
            root.gwritten = True
            if whole:
                return g, False
            return g.cols(m.col0, m.col0 + m.C), True
        g = root.gbuf
        if m.root is None:
            return g, True
        return g.cols(m.col0, m.col0 + m.C), True

    def grad_of(self, m: Mat) -> Optional[Mat]:
        root = m.root if m.root is not None else m
        if root.gbuf is None:
            return None
        return root.gbuf if m.root is None else root.gbuf.cols(m.col0, m.col0 + m.C)

    def set_grad(self, m: Mat, g: Mat):
        assert m.root is None
        m.gbuf = g
        m.gwritten = True

    # ---- kernel wrappers ----------------------------------------------------------------------------
    def ew(self, op, a: Optional[Mat], b: Optional[Mat], out: Mat, p=0.0, seed=0):
        ref = a if a is not None else b
        call("cris_elementwise", op, a.ptr if a else None, int(a.fp32) if a else 0, a.ld if a else 0,
             b.ptr if b else None, int(b.fp32) if b else 0, b.ld if b else 0, out.ptr, int(out.fp32), out.ld,
             ref.rows, ref.C, float(p), int(seed), self.seed_dev if p > 0 else None)

    def accumulate_into(self, src: Mat, dst_owner: Mat):
        """grad(dst_owner) (+)= src"""
        slot, acc = self.grad_slot(dst_owner)
        self.ew(0, src, slot if acc else None, slot)

    def gemm(self, A: Mat, B: Mat, D: Mat, M, N, K, *, a_mn=0, b_mn=0, batch=1, batch_inner=1, sA=(0, 0), sB=(0, 0),
             sD=(0, 0), alpha=1.0, bias=None, act=0, resid: Optional[Mat] = None, sR=(0, 0), mask_geom=None,
             colstats=None, tap_mode=0, taps=1, tap_off=None, b_tap_k=0, b_tap_n=0, d_tap_n=0, splits=1,
             accumulate=0, d_col_stride=0, a_rows=0, b_rows=0):
        g = GemmArgs()
        g.A, g.lda, g.strideA, g.strideA2 = A.ptr, A.ld, sA[0], sA[1]
        g.B, g.ldb, g.strideB, g.strideB2 = B.ptr, B.ld, sB[0], sB[1]
        g.D, g.ldd, g.strideD, g.strideD2 = D.ptr, D.ld, sD[0], sD[1]
        g.M, g.N, g.K, g.batch, g.batch_inner = M, N, K, batch, batch_inner
        g.a_mn, g.b_mn, g.d_fp32, g.accumulate = a_mn, b_mn, int(D.fp32), accumulate
        g.tap_mode, g.taps = tap_mode, taps
        if tap_off is not None:
            for i, v in enumerate(tap_off):
                g.tap_off[i] = v
        g.b_tap_k, g.b_tap_n, g.d_tap_n, g.splits = b_tap_k, b_tap_n, d_tap_n, splits
        g.alpha = alpha
        g.bias = bias
        g.act = act
        if resid is not None:
            g.resid, g.ldr, g.strideR, g.strideR2, g.resid_fp32 = resid.ptr, resid.ld, sR[0], sR[1], int(resid.fp32)
        if mask_geom is not None:
            g.mask_hp, g.mask_wp = mask_geom[1] + 2, mask_geom[2] + 2
        g.colstats = colstats
        g.a_rows, g.b_rows, g.d_col_stride = a_rows, b_rows, d_col_stride
        if _lib._prof is not None:
            nt = taps if tap_mode else 1
            cls = ("attention" if batch > 1 else "wgrad" if accumulate else "dgrad" if self.in_backward else "fwd")
            _lib.prof_tag = ("gemm_" + cls, 2.0 * M * N * K * nt * batch)
        if self.e.gemm_log is not None:
            self.e.gemm_log.append((M, N, K, batch, a_mn, b_mn, taps if tap_mode else 1, tap_mode, splits, int(D.fp32),
                                    int(resid is not None), int(colstats is not None)))
        gemm(g)

    @staticmethod
    def taps_of(geom):
        wp = geom[2] + 2
        return [dy * wp + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1)]

    def col_sum(self, m: Mat, width: int, hp=0, wp=0) -> torch.Tensor:
        """per-column sums of m[:, :width] (bias gradients); fp32 vector of the padded width."""
        wpad = _r8(width)
        out = self.f32(wpad)
        nb = max(1, min(592, m.rows // 64))
        for c0 in range(0, wpad, 2048):
            cw = min(2048, wpad - c0)
            pt = self.zeros_f32(min(nb, 64) * 2 * cw)
            call("cris_col_reduce", 2, m.ptr + c0 * m.esize, m.ld, int(m.fp32), None, 0, None, 0, None, 0, 0, None, None,
                 None, None, m.rows, cw, 0, hp, wp, pt.data_ptr(), nb)
            sm = self.f32(2 * cw)
            call("cris_bn_reduce_partials", pt.data_ptr(), min(nb, 64), cw, sm.data_ptr())
            out[c0:c0 + cw].copy_(sm[:cw])
        return out

    # ---- BatchNorm ---------------------------------------------------------------------------------
    def allreduce(self, t: torch.Tensor):
        """Sum the statistics vector over ranks: one peer-memory kernel on NVLink, else torch.distributed."""
        if self.xchg is not None and t.numel() <= PEER_SLOT_FLOATS and self.xslot < PEER_MAX_SLOTS:
            self.xchg.allreduce(self.xslot, t)
            self.xslot += 1
        else:
            if t.is_cuda and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("cross-rank BatchNorm statistics fell back to NCCL inside a CUDA-graph capture "
                                   f"(exchange site {self.xslot}, {t.numel()} floats)")
            dist.all_reduce(t)

    def fused_sync_ok(self, C: int) -> bool:
        return (self.xchg is not None and 2 * C <= PEER_SLOT_FLOATS
                and os.environ.get("CRIS_B200_FUSED_SYNCBN", "1") != "0")

    def bn_forward(self, z: Mat, prefix: str, relu: bool, resid: Optional[Mat] = None, out: Optional[Mat] = None,
                   partials=None, n_tiles=0) -> Mat:
        """BatchNorm over the rows of z (+ residual, ReLU).  Training = batch statistics (SyncBN-equivalent
        across ranks), eval = running statistics.  Reference: nn.BatchNorm2d/1d call sites in model/clip.py,
        model/layers.py; SyncBatchNorm (train.py:97-98)."""
        C = z.C
        gamma, beta = self.P[prefix + ".weight"], self.P[prefix + ".bias"]
        rm, rv = self.Bf[prefix + ".running_mean"], self.Bf[prefix + ".running_var"]
        coef = self.f32(4 * C)  # scale | shift | mean | invstd
        count_local = z.geom[0] * z.geom[1] * z.geom[2] if z.geom else z.rows
        count = float(count_local * (self.world if self.sync_bn else 1))
        sums = None
        if self.training:
            if partials is None:
                nb = max(1, min(592, z.rows // 64))
                n_tiles = min(nb, 64)
                partials = self.zeros_f32(n_tiles * 2 * C)
                call("cris_col_reduce", 0, z.ptr, z.ld, 0, None, 0, None, 0, None, 0, 0, None, None, None, None, z.rows,
                     C, 0, z.hp, z.wp, partials.data_ptr(), nb)
            sums = self.f32(2 * C)
            if self.sync_bn and self.fused_sync_ok(C):
                # ONE kernel per exchange site: partials -> local sums -> pushed to every peer over NVLink ->
                # rank-ordered global sums -> coefficients + running statistics (csrc/peer.cu peer_bn_sync_kernel)
                self.xchg.bn_sync_fwd(self.xslot, partials, n_tiles, C, count, gamma, beta, BN_EPS, BN_MOMENTUM, rm, rv, coef)
                self.xslot += 1
            elif self.sync_bn:
                call("cris_bn_reduce_partials", partials.data_ptr(), n_tiles, C, sums.data_ptr())
                self.allreduce(sums)
                call("cris_bn_coeffs", sums.data_ptr(), count, gamma.data_ptr(), beta.data_ptr(), BN_EPS, BN_MOMENTUM,
                     rm.data_ptr(), rv.data_ptr(), coef.data_ptr(), coef.data_ptr() + 4 * C, coef.data_ptr() + 8 * C,
                     coef.data_ptr() + 12 * C, C, 1)
            else:
                call("cris_bn_finalize_fwd", partials.data_ptr(), n_tiles, C, sums.data_ptr(), count, gamma.data_ptr(),
                     beta.data_ptr(), BN_EPS, BN_MOMENTUM, rm.data_ptr(), rv.data_ptr(), coef.data_ptr(),
                     coef.data_ptr() + 4 * C, coef.data_ptr() + 8 * C, coef.data_ptr() + 12 * C)
            nbt = self.Bf.get(prefix + ".num_batches_tracked")
            if nbt is not None:
                nbt.add_(1)
        else:
            call("cris_bn_coeffs", None, 1.0, gamma.data_ptr(), beta.data_ptr(), BN_EPS, BN_MOMENTUM, rm.data_ptr(),
                 rv.data_ptr(), coef.data_ptr(), coef.data_ptr() + 4 * C, coef.data_ptr() + 8 * C,
                 coef.data_ptr() + 12 * C, C, 0)
        y = out if out is not None else self.new(z.rows, C, False, z.geom)
        call("cris_bn_apply", z.ptr, z.ld, coef.data_ptr(), coef.data_ptr() + 4 * C, resid.ptr if resid else None,
             resid.ld if resid else 0, y.ptr, y.ld, z.rows, C, int(relu), z.hp, z.wp)

        def bwd():
            dy = self.grad_of(y)
            if dy is None:
                return
            nb = max(1, min(592, z.rows // 64))
            part = self.zeros_f32(min(nb, 64) * 2 * C)
            # without a residual the ReLU mask is recomputed from z (x*scale+shift > 0): y is not re-read
            ymask = y if (resid is not None or not relu) else None
            # residual + ReLU layers (bn3 of every bottleneck): the masked gradient dy*(y>0) is the residual branch's
            # gradient; the reduction pass writes it there and the apply pass reads it back instead of (dy, y)
            premasked = None
            if (relu and resid is not None and resid.need_grad and ymask is not None
                    and os.environ.get("CRIS_B200_BN_DZM", "1") != "0"):
                root = resid.root if resid.root is not None else resid
                stream_ok = (2048 <= z.rows < (1 << 31) and 8 <= C <= 2048 and (C & (C - 1)) == 0
                             and os.environ.get("CRIS_B200_BN_STREAM", "1") != "0"
                             and all(m.ptr % 16 == 0 and m.ld % 8 == 0 for m in (dy, ymask, z)))
                if stream_ok and root.gbuf is None and resid.root is None:
                    rslot, racc = self.grad_slot(resid)
                    call("cris_bn_bwd_reduce_masked", dy.ptr, dy.ld, ymask.ptr, ymask.ld, z.ptr, z.ld,
                         coef.data_ptr() + 8 * C, coef.data_ptr() + 12 * C, z.rows, C, z.hp, z.wp, rslot.ptr, rslot.ld,
                         part.data_ptr(), nb)
                    premasked = rslot
            if premasked is None:
                call("cris_col_reduce", 1, dy.ptr, dy.ld, 0, None, 0, ymask.ptr if ymask else None, ymask.ld if ymask else 0,
                     z.ptr, z.ld, 0, coef.data_ptr() + 8 * C, coef.data_ptr() + 12 * C, coef.data_ptr(),
                     coef.data_ptr() + 4 * C, z.rows, C, int(relu), z.hp, z.wp, part.data_ptr(), nb)
            bs = self.f32(2 * C)
            # parameter gradients are LOCAL sums (DDP averages them), dx needs the GLOBAL sums
            gb, gg = self.pg(prefix + ".bias"), self.pg(prefix + ".weight")
            if self.sync_bn and self.fused_sync_ok(C):
                self.xchg.bn_sync_bwd(self.xslot, part, min(nb, 64), C, bs, gb, gg)
                self.xslot += 1
            else:
                call("cris_stats_finalize_bwd", part.data_ptr(), min(nb, 64), C, bs.data_ptr(), gb.data_ptr(), gg.data_ptr())
                if self.sync_bn:
                    self.allreduce(bs)
            dz = self.new(z.rows, C, False, z.geom)
            if premasked is not None:
                call("cris_bn_bwd_apply", premasked.ptr, premasked.ld, None, 0, z.ptr, z.ld, coef.data_ptr() + 8 * C,
                     coef.data_ptr() + 12 * C, gamma.data_ptr(), beta.data_ptr(), bs.data_ptr(), count, dz.ptr, dz.ld,
                     None, 0, 0, z.rows, C, 0, z.hp, z.wp)
                self.set_grad(z, dz)
                return
            dres_ptr, dres_ld, dres_acc = None, 0, 0
            if resid is not None and resid.need_grad:
                slot, acc = self.grad_slot(resid)
                dres_ptr, dres_ld, dres_acc = slot.ptr, slot.ld, int(acc)
            call("cris_bn_bwd_apply", dy.ptr, dy.ld, ymask.ptr if ymask else None, ymask.ld if ymask else 0, z.ptr, z.ld,
                 coef.data_ptr() + 8 * C, coef.data_ptr() + 12 * C, gamma.data_ptr(), beta.data_ptr(), bs.data_ptr(),
                 count, dz.ptr, dz.ld, dres_ptr, dres_ld, dres_acc, z.rows, C, int(relu), z.hp, z.wp)
            self.set_grad(z, dz)

        if self.training:
            self.on_backward(bwd)
        return y

    # ---- convolution (implicit GEMM on tcgen05) ---------------------------------------------------
    def conv(self, x: Mat, wname: str, k: int, stats: bool, bias_name: Optional[str] = None,
             cin: Optional[int] = None, as_matrix: bool = False):
        """z = conv_kxk(x) (stride 1, zero padding k//2), optional bias; returns (z, colstats partials, tiles).
        as_matrix: the weight is used as a plain [Cout, Cin*kh*kw] matrix over pre-gathered patches (stem)."""
        wp = self.w(wname, as_matrix)
        Wt = self.P[wname]
        cout = Wt.shape[0]
        w_cols = Wt.numel() // cout if as_matrix else Wt.shape[1]
        cin = w_cols if cin is None else cin
        cin_pad = _r8(w_cols)
        z = self.new(x.rows, cout, False, x.geom)
        n_tiles = min(64, (x.rows + 127) // 128)
        part = self.f32(n_tiles * 2 * cout, zero=True) if stats else None
        bias = self.P[bias_name].data_ptr() if bias_name else None
        offs = self.taps_of(x.geom) if k == 3 else None
        probe = self.e.probe_name == wname
        if probe:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if (k == 3 and bias is None and cin in (32, 64) and cout in (32, 64) and x.geom is not None and not as_matrix
                and os.environ.get("CRIS_B200_HALO_CONV", "1") == "1"):
            # small-channel path (csrc/conv_halo.cu): one halo tile per 128*SUB output rows, the nine taps are
            # descriptor offsets into it instead of nine L2 re-reads (measured: -1.1 ms per step, profiles/r02_flag_sweep.txt;
            # CRIS_B200_HALO_CONV=0 selects the generic implicit GEMM)
            N_, H_, W_ = x.geom
            call("cris_conv3x3_halo", x.ptr, x.ld, wp.ptr, wp.ld, cin_pad, z.ptr, z.ld,
                 part.data_ptr() if stats else None, N_, H_, W_, cin, cout)
        else:
            self.gemm(x, wp, z, x.rows, cout, cin, bias=bias, mask_geom=x.geom,
                      colstats=part.data_ptr() if stats else None, tap_mode=TAP_ACCUM if k == 3 else TAP_NONE,
                      taps=9 if k == 3 else 1, tap_off=offs, b_tap_k=cin_pad if k == 3 else 0)
        if probe:
            ev1.record()
            self.e.probe_events.append((ev0, ev1))

        def bwd():
            dz = self.grad_of(z)
            if dz is None:
                return
            def wgrad():
                if bias_name:
                    self.pg(bias_name).copy_(self.col_sum(dz, cout, z.hp, z.wp)[:cout])
                # wgrad: dW[co][ci][tap] += sum_rows dz[row][co] * x[row + off_tap][ci]   (fp32, split-K atomics)
                gw = self.pg(wname)
                gwm = Mat(gw, cout, w_cols * (9 if k == 3 else 1), fp32=True)
                # 0 = the library plans tile width and split-K together (whole waves of its persistent grid; partial
                # sums meet in fp32 TMA reductions in arrival order).  CRIS_B200_DETERMINISTIC_WGRAD=1: one unit per
                # output tile (splits = 1) -> every gradient element is written by exactly one reduction into the
                # zeroed buffer, i.e. bitwise repeatable weight gradients (slower: small layers no longer fill the
                # machine).
                splits = 1 if os.environ.get("CRIS_B200_DETERMINISTIC_WGRAD", "0") == "1" else 0
                if k == 3:
                    # the nine taps accumulate into a zeroed [cout][9][cin_pad] fp32 scratch with unit column stride
                    # (TMA reductions, csrc/gemm_tc.cu EPI_ACCUM), then one small kernel writes the reference's OIHW
                    acc = Mat(self.zeros_f32(cout * 9 * cin_pad), cout, 9 * cin_pad, fp32=True)
                    self.gemm(dz, x, acc, cout, cin, x.rows, a_mn=1, b_mn=1, tap_mode=TAP_WGRAD, taps=9, tap_off=offs,
                              d_tap_n=cin_pad, splits=splits, accumulate=1)
                    call("cris_unpack_conv_wgrad", acc.ptr, gw.data_ptr(), cout, w_cols, 9, cin_pad)
                else:
                    self.gemm(dz, x, gwm, cout, cin, x.rows, a_mn=1, b_mn=1, splits=splits, accumulate=1)

            self.leaf_branch(wgrad, (dz, x, z))
            # dgrad: dx[row] = sum_tap dz[row - off_tap] * W_tap
            if x.need_grad:
                slot, acc = self.grad_slot(x)
                n_in = min(cin, x.C)
                if (k == 3 and not acc and n_in == cin and cin in (32, 64) and cout in (32, 64) and x.geom is not None
                        and not as_matrix and os.environ.get("CRIS_B200_HALO_CONV", "1") == "1"):
                    # dx = conv3x3(dz, mirrored / transposed weights) through the halo-tile kernel
                    cop = _r8(cout)
                    wt = self.new(cin, 9 * cop)
                    call("cris_pack_conv_weight_dgrad", self.P[wname].data_ptr(), wt.ptr, cout, cin, cop)
                    N_, H_, W_ = x.geom
                    call("cris_conv3x3_halo", dz.ptr, dz.ld, wt.ptr, wt.ld, cop, slot.ptr, slot.ld, None, N_, H_, W_,
                         cout, cin)
                    return
                self.gemm(dz, wp, slot, x.rows, n_in, cout, b_mn=1, mask_geom=x.geom, resid=slot if acc else None,
                          tap_mode=TAP_ACCUM if k == 3 else TAP_NONE, taps=9 if k == 3 else 1,
                          tap_off=[-o for o in offs] if k == 3 else None, b_tap_n=cin_pad if k == 3 else 0,
                          b_rows=cout)

        if self.training:
            self.on_backward(bwd)
        return z, part, n_tiles

    def conv_bn(self, x: Mat, conv_name: str, bn_prefix: str, k: int, relu=True, resid=None, out=None, cin=None):
        if not self.training and self.e.fold_eval_bn:
            return self.conv_bn_folded(x, conv_name, bn_prefix, k, relu, resid, out, cin)
        z, part, nt = self.conv(x, conv_name, k, stats=self.training, cin=cin)
        return self.bn_forward(z, bn_prefix, relu, resid, out, part, nt)

    def conv_bn_folded(self, x: Mat, conv_name: str, bn_prefix: str, k: int, relu, resid, out, cin) -> Mat:
        """model.eval(): y = relu?(conv(x) * scale + shift (+ resid)) with scale = gamma / sqrt(running_var + eps),
        shift = beta - running_mean * scale — the scale is folded into the bf16 weights, the shift is the GEMM's
        bias, residual add and ReLU run in the GEMM epilogue: no BatchNorm pass over the activations at all
        (model/layers.py:8-11, model/clip.py:44-57 under model.eval(); engine/engine.py:100,171)."""
        Wt = self.P[conv_name]
        cout, w_cols = Wt.shape[0], Wt.shape[1]
        cin = w_cols if cin is None else cin
        cin_pad = _r8(w_cols)
        gamma, beta = self.P[bn_prefix + ".weight"], self.P[bn_prefix + ".bias"]
        rm, rv = self.Bf[bn_prefix + ".running_mean"], self.Bf[bn_prefix + ".running_var"]
        coef = self.e.eval_coef(bn_prefix, cout, self.dev)  # scale | shift | mean | invstd, persistent per layer
        if bn_prefix not in getattr(self, "bn_done", ()):
            call("cris_bn_coeffs", None, 1.0, gamma.data_ptr(), beta.data_ptr(), BN_EPS, BN_MOMENTUM, rm.data_ptr(),
                 rv.data_ptr(), coef.data_ptr(), coef.data_ptr() + 4 * cout, coef.data_ptr() + 8 * cout,
                 coef.data_ptr() + 12 * cout, cout, 0)
        wp = self.e.packed.get_folded(conv_name, Wt, coef.data_ptr())
        y = out if out is not None else self.new(x.rows, cout, False, x.geom)
        act = ACT_NONE if not relu else (ACT_RELU_POST if resid is not None else ACT_RELU)
        self.gemm(x, wp, y, x.rows, cout, cin, bias=coef.data_ptr() + 4 * cout, act=act, resid=resid,
                  mask_geom=x.geom, tap_mode=TAP_ACCUM if k == 3 else TAP_NONE, taps=9 if k == 3 else 1,
                  tap_off=self.taps_of(x.geom) if k == 3 else None, b_tap_k=cin_pad if k == 3 else 0)
        return y

    # ---- resampling -------------------------------------------------------------------------------
    def avgpool(self, x: Mat, out: Optional[Mat] = None) -> Mat:
        N, H, W = x.geom
        y = out if out is not None else self.padded(N, H // 2, W // 2, x.C)
        y.geom = (N, H // 2, W // 2)
        call("cris_avgpool2_fwd", x.ptr, x.ld, y.ptr, y.ld, N, H, W, x.C)

        def bwd():
            dy = self.grad_of(y)
            if dy is None or not x.need_grad:
                return
            slot, acc = self.grad_slot(x)
            call("cris_avgpool2_bwd", dy.ptr, dy.ld, slot.ptr, slot.ld, int(acc), N, H, W, x.C)

        if self.training:
            self.on_backward(bwd)
        return y

    def upsample(self, x: Mat, out: Optional[Mat] = None) -> Mat:
        N, H, W = x.geom
        y = out if out is not None else self.padded(N, 2 * H, 2 * W, x.C)
        call("cris_upsample2x_fwd", x.ptr, x.ld, y.ptr, y.ld, N, H, W, x.C)

        def bwd():
            dy = self.grad_of(y)
            if dy is None or not x.need_grad:
                return
            slot, acc = self.grad_slot(x)
            call("cris_upsample2x_bwd", dy.ptr, dy.ld, slot.ptr, slot.ld, int(acc), N, H, W, x.C)

        if self.training:
            self.on_backward(bwd)
        return y

    # ---- token-side building blocks -----------------------------------------------------------------
    def linear(self, x: Mat, wname: str, bname: Optional[str], *, act=ACT_NONE, out_fp32=False, w_rows=None,
               stats=False, transposed_weight=False, out: Optional[Mat] = None):
        """y = act(x @ W^T + b) (nn.Linear / 1x1 projections of MHA).  `w_rows=(r0, r1)` uses a row slice of the
        weight (packed in_proj of nn.MultiheadAttention).  transposed_weight: W is stored [in, out]
        (CLIP text_projection, model/clip.py:451-452)."""
        Wt = self.P[wname]
        wp = self.w(wname)
        if transposed_weight:
            n_in, n_out = Wt.shape
            r0, r1 = 0, n_out
            wv = wp
        else:
            r0, r1 = (0, Wt.shape[0]) if w_rows is None else w_rows
            n_out, n_in = r1 - r0, Wt.shape[1]
            wv = wp.rows_slice(r0, r1)
        bias_ptr = (self.P[bname].data_ptr() + 4 * r0) if bname else None
        y = out if out is not None else self.new(x.rows, n_out, out_fp32, None, ld=_r8(n_out))
        n_tiles = min(64, (x.rows + 127) // 128)
        part = self.f32(n_tiles * 2 * n_out, zero=True) if stats else None
        self.gemm(x, wv, y, x.rows, n_out, n_in, b_mn=1 if transposed_weight else 0, bias=bias_ptr, act=act,
                  colstats=part.data_ptr() if stats else None, b_rows=n_in if transposed_weight else 0)

        def bwd():
            dy = self.grad_of(y)
            if dy is None:
                return
            if act == ACT_RELU:
                t = self.new(y.rows, y.C, False, None, ld=y.ld)
                self.ew(5, y, dy, t)
                dy = t
            elif dy.fp32:
                # bf16 copy for the tensor-core GEMMs; odd widths (proj.txt: 9C+1) are cast over the padded pitch
                wpad = _r8(y.C)
                t = self.new(y.rows, y.C, False, None, ld=wpad)
                self.ew(0, Mat(dy.buf, dy.rows, wpad, dy.ld, True, ptr=dy.ptr), None, Mat(t.buf, t.rows, wpad, wpad))
                dy = t
            dy_ = dy

            def wgrad(dy=dy_):
                self._linear_wgrad(dy, x, wname, bname, r0, r1, n_in, n_out, transposed_weight)

            self.leaf_branch(wgrad, (dy_, x, y))
            if x.need_grad:
                slot, acc = self.grad_slot(x)
                if slot.fp32:
                    raise RuntimeError("linear input gradients are bf16")
                self.gemm(dy, wv, slot, x.rows, n_in, n_out, b_mn=0 if transposed_weight else 1,
                          resid=slot if acc else None, b_rows=0 if transposed_weight else n_out)

        if self.training:
            self.on_backward(bwd)
        return (y, part, n_tiles) if stats else y

    def _linear_wgrad(self, dy, x, wname, bname, r0, r1, n_in, n_out, transposed_weight):
        if True:
            if bname:
                if os.environ.get("CRIS_B200_BIAS_MMA", "1") == "1" and dy.rows >= 1024 and n_out % 8 == 0:
                    # bias gradient on the tensor cores: 1^T . dy as an M = 1 GEMM (the ones vector is the only
                    # in-bounds row of the A box, TMA zero-fills the other 127); split-K atomics land directly in
                    # the zero-filled gradient buffer: one launch instead of reduce + finalize + two copies
                    kp = _r8(dy.rows)
                    ones = self.e.const(("ones_bf16", kp), lambda: torch.ones(1, kp, dtype=torch.bfloat16), self.dev)
                    gbv = self.pg(bname)
                    dst = Mat(gbv, 1, n_out, fp32=True, ptr=gbv.data_ptr() + 4 * r0)
                    self.gemm(Mat(ones, 1, dy.rows, ld=kp), dy, dst, 1, n_out, dy.rows, b_mn=1, splits=0, accumulate=1)
                else:
                    sm = self.col_sum(dy, n_out)
                    self.pg(bname)[r0:r1].copy_(sm[:n_out])
            gw = self.pg(wname)
            det = 1 if os.environ.get("CRIS_B200_DETERMINISTIC_WGRAD", "0") == "1" else 0
            if transposed_weight:
                gwm = Mat(gw, n_in, n_out, fp32=True)
                sp = det
                self.gemm(x, dy, gwm, n_in, n_out, x.rows, a_mn=1, b_mn=1, splits=sp, accumulate=1)
            else:
                gwm = Mat(gw, n_out, n_in, fp32=True, ptr=gw.data_ptr() + 4 * r0 * n_in)
                sp = det
                self.gemm(dy, x, gwm, n_out, n_in, x.rows, a_mn=1, b_mn=1, splits=sp, accumulate=1)

    def layernorm(self, x: Mat, prefix: str, add: Optional[torch.Tensor] = None, want_y=True, want_y2=False):
        """y = LN(x) (bf16); y2 = y + add[row % period] (bf16) — nn.LayerNorm (+ with_pos_embed,
        model/layers.py:221-222)."""
        C = x.C
        gamma, beta = self.P[prefix + ".weight"], self.P[prefix + ".bias"]
        y = self.new(x.rows, C) if want_y else None
        y2 = self.new(x.rows, C) if want_y2 else None
        stats = self.f32(2 * x.rows)
        call("cris_layernorm_fwd", x.ptr, int(x.fp32), x.ld, gamma.data_ptr(), beta.data_ptr(),
             add.data_ptr() if add is not None else None, C, add.shape[0] if add is not None else 1,
             y.ptr if y else None, 0, C, y2.ptr if y2 else None, C, stats.data_ptr(), stats.data_ptr() + 4 * x.rows,
             x.rows, C, LN_EPS)

        def bwd():
            d1 = self.grad_of(y) if y is not None else None
            d2 = self.grad_of(y2) if y2 is not None else None
            if d1 is None and d2 is None:
                return
            if d1 is None:
                d1, d2 = d2, None
            # one pass: dx and the gamma/beta gradients (accumulated into the zero-filled parameter-gradient buffer)
            gb, gg = self.pg(prefix + ".bias"), self.pg(prefix + ".weight")
            slot, acc = self.grad_slot(x) if x.need_grad else (None, False)
            call("cris_layernorm_bwd", d1.ptr, int(d1.fp32), d1.ld, d2.ptr if d2 else None, d2.ld if d2 else 0,
                 x.ptr, int(x.fp32), x.ld, gamma.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * x.rows,
                 slot.ptr if slot else None, int(slot.fp32) if slot else 0, slot.ld if slot else 0, int(acc),
                 gg.data_ptr(), gb.data_ptr(), x.rows, C)

        if self.training:
            self.on_backward(bwd)
        return y, y2

    def residual_add(self, x: Mat, h: Mat, p: float) -> Mat:
        """x_new = x + dropout(h) on the fp32 residual stream (model/layers.py:237,245,249; clip.py:263-264)."""
        out = self.new(x.rows, x.C, True)
        sd = self.seed() if p > 0 else 0
        self.ew(1, x, h, out, p, sd)

        def bwd():
            g = self.grad_of(out)
            if g is None:
                return
            # identity branch: grad(x) (+)= g ; dropout branch: grad(h) (+)= dropout_mask(g)
            hroot = h.root if h.root is not None else h
            if p == 0 and hroot.gbuf is None and h.root is None and g.ld == h.C:
                # no dropout: grad(h) IS grad(out) (fp32); its consumers (LayerNorm / Linear backward) accept fp32
                self.set_grad(h, g)
            else:
                slot, acc = self.grad_slot(h)
                if acc:
                    t = self.new(h.rows, h.C, slot.fp32)
                    self.ew(2, g, None, t, p, sd)
                    self.ew(0, slot, t, slot)
                else:
                    self.ew(2, g, None, slot, p, sd)
            root = x.root if x.root is not None else x
            if root.gbuf is None and x.root is None and g.ld == x.C and g.fp32 == x.fp32:
                self.set_grad(x, g)  # first contribution: alias grad(out) instead of copying it
            else:
                slot, acc = self.grad_slot(x)
                self.ew(0, g, slot if acc else None, slot)

        if self.training:
            self.on_backward(bwd)
        return out

    def attention(self, q: Mat, k: Mat, v: Mat, B, heads, Lq, Lk, causal=False, key_pad=False, p_drop=0.0) -> Mat:
        """softmax(q k^T / sqrt(64) + masks) v per head (core of F.multi_head_attention_forward as called at
        model/clip.py:119-139,255-260 and model/layers.py:235,240-243).  q,k,v: [B*L, heads*64] bf16 (possibly
        column slices of a packed projection)."""
        E = heads * 64
        if (Lk >= 64 and not causal and not key_pad and os.environ.get("CRIS_B200_FUSED_ATTN", "1") != "0"):
            return self.attention_fused(q, k, v, B, heads, Lq, Lk, p_drop)
        Lkp = _r8(Lk)
        nb = B * heads
        S = self.new(nb * Lq, Lk, False, None, ld=Lkp)
        sS = (heads * Lq * Lkp, Lq * Lkp)
        alpha = 1.0 / 8.0
        self.gemm(q, k, S, Lq, Lk, 64, batch=nb, batch_inner=heads, sA=(Lq * q.ld, 64), sB=(Lk * k.ld, 64), sD=sS,
                  alpha=alpha)
        Pd = self.new(nb * Lq, Lk, False, None, ld=Lkp) if p_drop > 0 else None
        sd = self.seed() if p_drop > 0 else 0
        call("cris_softmax_fwd", S.ptr, S.ptr, Pd.ptr if Pd else None, Lkp, Lq * Lkp, nb, Lq, Lk, heads,
             self.word.data_ptr() if key_pad else None, int(causal), float(p_drop), int(sd),
             self.seed_dev if p_drop > 0 else None)
        Puse = Pd if Pd is not None else S
        o = self.new(B * Lq, E)
        self.gemm(Puse, v, o, Lq, 64, Lk, b_mn=1, batch=nb, batch_inner=heads, sA=sS, sB=(Lk * v.ld, 64),
                  sD=(Lq * E, 64), b_rows=Lk)

        def bwd():
            do = self.grad_of(o)
            if do is None:
                return
            dP = self.new(nb * Lq, Lk, False, None, ld=Lkp)
            self.gemm(do, v, dP, Lq, Lk, 64, batch=nb, batch_inner=heads, sA=(Lq * do.ld, 64), sB=(Lk * v.ld, 64),
                      sD=sS)
            # dV = Pd^T do
            slot, acc = self.grad_slot(v)
            self.gemm(Puse, do, slot, Lk, 64, Lq, a_mn=1, b_mn=1, batch=nb, batch_inner=heads, sA=sS,
                      sB=(Lq * do.ld, 64), sD=(Lk * slot.ld, 64), resid=slot if acc else None,
                      sR=(Lk * slot.ld, 64), a_rows=Lq, b_rows=Lq)
            call("cris_softmax_bwd", S.ptr, dP.ptr, Lkp, Lq * Lkp, nb, Lq, Lk, float(p_drop), int(sd),
                 self.seed_dev if p_drop > 0 else None)
            slot, acc = self.grad_slot(q)
            self.gemm(dP, k, slot, Lq, 64, Lk, b_mn=1, batch=nb, batch_inner=heads, sA=sS, sB=(Lk * k.ld, 64),
                      sD=(Lq * slot.ld, 64), alpha=alpha, resid=slot if acc else None, sR=(Lq * slot.ld, 64),
                      b_rows=Lk)
            slot, acc = self.grad_slot(k)
            self.gemm(dP, q, slot, Lk, 64, Lq, a_mn=1, b_mn=1, batch=nb, batch_inner=heads, sA=sS,
                      sB=(Lq * q.ld, 64), sD=(Lk * slot.ld, 64), alpha=alpha, resid=slot if acc else None,
                      sR=(Lk * slot.ld, 64), a_rows=Lq, b_rows=Lq)

        if self.training:
            self.on_backward(bwd)
        return o

    def attention_fused(self, q: Mat, k: Mat, v: Mat, B, heads, Lq, Lk, p_drop=0.0) -> Mat:
        """The same attention with the scores kept on the SM (csrc/attention.cu): one forward kernel, and a backward
        that recomputes P from Q, K and the saved row log-sum-exp — S, P, dropout(P), dP never exist in HBM."""
        E = heads * 64
        o = self.new(B * Lq, E)
        lse = self.f32(B * heads * Lq)
        sd = self.seed() if p_drop > 0 else 0
        alpha = 1.0 / 8.0
        call("cris_attention_fwd", q.ptr, q.ld, k.ptr, k.ld, v.ptr, v.ld, o.ptr, o.ld, lse.data_ptr(), B, heads, Lq, Lk,
             alpha, float(p_drop), int(sd), self.seed_dev if p_drop > 0 else None)

        def bwd():
            do = self.grad_of(o)
            if do is None:
                return
            dq_acc = Mat(self.zeros_f32(B * Lq * E), B * Lq, E, fp32=True)
            dk, dv = self.new(B * Lk, E), self.new(B * Lk, E)
            dsum = self.f32(B * heads * Lq)
            call("cris_attention_bwd", q.ptr, q.ld, k.ptr, k.ld, v.ptr, v.ld, o.ptr, o.ld, do.ptr, do.ld, lse.data_ptr(),
                 dsum.data_ptr(), dq_acc.ptr, dq_acc.ld, dk.ptr, dk.ld, dv.ptr, dv.ld, B, heads, Lq, Lk, alpha,
                 float(p_drop), int(sd), self.seed_dev if p_drop > 0 else None)
            for src, owner in ((dq_acc, q), (dk, k), (dv, v)):
                slot, acc = self.grad_slot(owner)
                self.ew(0, src, slot if acc else None, slot)

        if self.training:
            self.on_backward(bwd)
        return o

    # =================================================================================================
    # the model
    # =================================================================================================
    def forward(self):
        if self.training:
            self.e.step_counter(self.dev).add_(7919)
        self.bn_done = set()
        if self.e.packed.force and os.environ.get("CRIS_B200_PACK_MULTI", "1") != "0":
            if not self.training and self.e._bn_multi is not None:
                # eval graph: the coefficients of every folded BatchNorm in one launch, then (below) every weight copy
                _, tab, n, blocks, names = self.e._bn_multi
                call("cris_bn_coeffs_multi", tab.data_ptr(), n, blocks, BN_EPS)
                self.bn_done = names
            self.e.packed.refresh_all()  # graph capture: all bf16 weight copies refreshed by one launch
        # The text tower (12 blocks on 17 tokens: ~250 latency-bound launches forward + backward) is independent of
        # the image encoder until the FPN.  Inside a CUDA-graph capture it is issued on a second stream, i.e. as a
        # parallel branch of the graph: its tiny kernels fill the tails of the image encoder's kernels instead of
        # adding ~4 ms of mostly idle GPU per step.  (Eager launches keep one stream: host-bound anyway, and the
        # caching allocator would need cross-stream bookkeeping.)
        self.text_range = None
        if (self.dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
                and os.environ.get("CRIS_B200_TEXT_STREAM", "1") != "0"
                and int(os.environ.get("CRIS_B200_BWD_SEGMENTS", "1")) <= 1):
            main = torch.cuda.current_stream(self.dev)
            side = self.e.side_stream(self.dev)
            side.wait_stream(main)
            t0 = len(self.tape)
            with torch.cuda.stream(side):
                wfeat, state = self.encode_text()
            self.text_range = (t0, len(self.tape))
            c3, c4, c5 = self.encode_image()
            self.image_end = len(self.tape)
            main.wait_stream(side)
        else:
            c3, c4, c5 = self.encode_image()
            wfeat, state = self.encode_text()
        fq = self.fpn(c3, c4, c5, state)
        fq = self.decoder(fq, wfeat)
        self.projector_and_loss(fq, state)

    # ---- image encoder: model/clip.py:207-223 ------------------------------------------------------------
    def encode_image(self):
        v = "backbone.visual"
        B, _, Hin, Win = self.img.shape
        if Hin % 32 or Win % 32:
            raise ValueError("image size must be a multiple of 32")
        w1 = self.P[v + ".conv1.weight"]
        # stem conv1 (3x3, stride 2) = im2col (27 taps, padded to 32) + the tcgen05 GEMM core (model/clip.py:165-170)
        patches = self.padded(B, Hin // 2, Win // 2, 32)
        call("cris_stem_im2col", self.img.data_ptr(), patches.ptr, B, Hin, Win)
        patches.need_grad = False
        z, part, nt = self.conv(patches, v + ".conv1.weight", 1, stats=self.training, cin=w1.numel() // w1.shape[0],
                                as_matrix=True)
        x = self.bn_forward(z, v + ".bn1", True, partials=part, n_tiles=nt)
        x = self.conv_bn(x, v + ".conv2.weight", v + ".bn2", 3)
        x = self.conv_bn(x, v + ".conv3.weight", v + ".bn3", 3)
        x = self.avgpool(x)
        self.tap("stem", x)
        feats = []
        vis = self.model.backbone.visual
        for li in (1, 2, 3, 4):
            layer = getattr(vis, f"layer{li}")
            for bi, blk in enumerate(layer):
                x = self.bottleneck(x, f"{v}.layer{li}.{bi}", blk.stride, blk.downsample is not None)
            feats.append(x)
            self.tap(f"layer{li}", x)
        c5 = self.attnpool(feats[3], v + ".attnpool", vis.attnpool.num_heads)
        self.tap("attnpool", c5)
        return feats[1], feats[2], c5

    def bottleneck(self, x: Mat, p: str, stride: int, has_down: bool) -> Mat:
        """model/clip.py:44-57."""
        o = self.conv_bn(x, p + ".conv1.weight", p + ".bn1", 1)
        o = self.conv_bn(o, p + ".conv2.weight", p + ".bn2", 3)
        if stride > 1:
            o = self.avgpool(o)
        if has_down:
            idn = self.avgpool(x) if stride > 1 else x
            idn = self.conv_bn(idn, p + ".downsample.0.weight", p + ".downsample.1", 1, relu=False)
        else:
            idn = x
        return self.conv_bn(o, p + ".conv3.weight", p + ".bn3", 1, relu=True, resid=idn)

    def attnpool(self, x: Mat, p: str, heads: int) -> Mat:
        """model/clip.py:110-144."""
        B, H, W = x.geom
        E = x.C
        T = H * W
        zc, part, nt = self.conv(x, p + ".connect.0.weight", 1, stats=self.training)
        pos = self.P[p + ".positional_embedding"]
        sp = int(round(math.sqrt(pos.shape[0] - 1)))
        R = self.e.const(("bicubic", sp, H, W), lambda: _bicubic_matrix(sp, H, W), self.dev)
        posr = self.f32(T * E)
        call("cris_small_matmul", R.data_ptr(), pos.data_ptr() + 4 * E, posr.data_ptr(), T, sp * sp, E, 0, 0)
        tok = self.new(B * T, E)
        call("cris_padded_to_tokens", x.ptr, x.ld, posr.data_ptr(), E, tok.ptr, 0, tok.ld, B, H, W, E)

        def bwd_tok():
            dt = self.grad_of(tok)
            if dt is None:
                return
            dposr = self.f32(T * E)
            call("cris_batch_reduce", dt.ptr, 0, dt.ld, dposr.data_ptr(), E, B, T, E, 0)
            gp = self.pg(p + ".positional_embedding")
            call("cris_small_matmul", R.data_ptr(), dposr.data_ptr(), gp.data_ptr() + 4 * E, T, sp * sp, E, 1, 0)
            slot, acc = self.grad_slot(x)
            if acc:
                t = self.padded(B, H, W, E)
                call("cris_tokens_to_padded", dt.ptr, 0, dt.ld, t.ptr, t.ld, B, H, W, E)
                self.ew(0, slot, t, slot)
            else:
                call("cris_tokens_to_padded", dt.ptr, 0, dt.ld, slot.ptr, slot.ld, B, H, W, E)

        if self.training:
            self.on_backward(bwd_tok)
        # k_proj, q_proj, v_proj are adjacent Linear layers -> three GEMMs over the same token matrix
        qkv = self.new(B * T, 3 * E)
        k = self.linear(tok, p + ".k_proj.weight", p + ".k_proj.bias", out=qkv.cols(0, E))
        q = self.linear(tok, p + ".q_proj.weight", p + ".q_proj.bias", out=qkv.cols(E, 2 * E))
        vv = self.linear(tok, p + ".v_proj.weight", p + ".v_proj.bias", out=qkv.cols(2 * E, 3 * E))
        o = self.attention(q, k, vv, B, heads, T, T)
        o = self.linear(o, p + ".c_proj.weight", p + ".c_proj.bias")
        Co = o.C
        op = self.padded(B, H, W, Co)
        call("cris_tokens_to_padded", o.ptr, 0, o.ld, op.ptr, op.ld, B, H, W, Co)

        def bwd_op():
            g = self.grad_of(op)
            if g is None:
                return
            slot, acc = self.grad_slot(o)
            assert not acc
            call("cris_padded_to_tokens", g.ptr, g.ld, None, 0, slot.ptr, 0, slot.ld, B, H, W, Co)

        if self.training:
            self.on_backward(bwd_op)
        return self.bn_forward(zc, p + ".connect.1", True, resid=op, partials=part, n_tiles=nt)

    # ---- text encoder: model/clip.py:439-456 ---------------------------------------------------------------
    def encode_text(self):
        b = "backbone"
        B, L = self.word.shape
        table = self.P[b + ".token_embedding.weight"]
        pos = self.P[b + ".positional_embedding"]
        E = table.shape[1]
        heads = E // 64
        x = self.new(B * L, E, True)
        call("cris_embed_fwd", self.word.data_ptr(), table.data_ptr(), pos.data_ptr(), x.ptr, B, L, E)
        x_embed = x  # `x` is rebound by the residual blocks below; closures must capture the embedding output

        def bwd_embed():
            g = self.grad_of(x_embed)
            if g is None:
                return
            call("cris_embed_bwd", self.word.data_ptr(), g.ptr, self.pg(b + ".token_embedding.weight").data_ptr(),
                 self.pg(b + ".positional_embedding").data_ptr(), B, L, E)

        if self.training:
            self.on_backward(bwd_embed)
        nblk = len(self.model.backbone.transformer.resblocks)
        for i in range(nblk):
            p = f"{b}.transformer.resblocks.{i}"
            h, _ = self.layernorm(x, p + ".ln_1")
            qkv = self.linear(h, p + ".attn.in_proj_weight", p + ".attn.in_proj_bias")
            o = self.attention(qkv.cols(0, E), qkv.cols(E, 2 * E), qkv.cols(2 * E, 3 * E), B, heads, L, L, causal=True)
            o = self.linear(o, p + ".attn.out_proj.weight", p + ".attn.out_proj.bias")
            x = self.residual_add(x, o, 0.0)
            h, _ = self.layernorm(x, p + ".ln_2")
            m = self.linear(h, p + ".mlp.c_fc.weight", p + ".mlp.c_fc.bias")
            a = self.quickgelu(m)
            o = self.linear(a, p + ".mlp.c_proj.weight", p + ".mlp.c_proj.bias")
            x = self.residual_add(x, o, 0.0)
        wfeat, _ = self.layernorm(x, b + ".ln_final")
        sin = self.new(B, E)
        call("cris_eot_gather", self.word.data_ptr(), wfeat.ptr, 0, wfeat.ld, sin.ptr, sin.ld, B, L, E)

        def bwd_eot():
            g = self.grad_of(sin)
            if g is None:
                return
            slot, acc = self.grad_slot(wfeat)
            if not acc:
                slot.buf.zero_()
            call("cris_eot_scatter", self.word.data_ptr(), g.ptr, 0, g.ld, slot.ptr, int(slot.fp32), slot.ld, B, L, E)

        if self.training:
            self.on_backward(bwd_eot)
        state = self.linear(sin, b + ".text_projection", None, transposed_weight=True)
        self.tap("word", wfeat)
        self.tap("state", state)
        return wfeat, state

    def quickgelu(self, m: Mat) -> Mat:
        a = self.new(m.rows, m.C, False, None, ld=m.ld)
        self.ew(3, m, None, a)

        def bwd():
            g = self.grad_of(a)
            if g is None:
                return
            slot, acc = self.grad_slot(m)
            assert not acc
            self.ew(4, m, g, slot)

        if self.training:
            self.on_backward(bwd)
        return a

    # ---- FPN neck: model/layers.py:282-309 --------------------------------------------------------------------
    def fpn(self, c3: Mat, c4: Mat, c5: Mat, state: Mat) -> Mat:
        B = state.rows
        n = "neck"
        zs, part, nt = self.linear(state, n + ".txt_proj.0.weight", None, stats=True)
        if not self.training:
            part = None
        s = self.bn_forward(zs, n + ".txt_proj.1", True, partials=part if self.training else None, n_tiles=nt)
        f5a = self.conv_bn(c5, n + ".f1_v_proj.0.weight", n + ".f1_v_proj.1", 1)
        N5, H5, W5 = f5a.geom
        g = self.new(f5a.rows, f5a.C, False, f5a.geom)
        rpi = (H5 + 2) * (W5 + 2)
        call("cris_mul_bcast", f5a.ptr, f5a.ld, s.ptr, s.ld, g.ptr, g.ld, f5a.rows, rpi, f5a.C)

        def bwd_gate():
            dg = self.grad_of(g)
            if dg is None:
                return
            slot, acc = self.grad_slot(f5a)
            assert not acc
            call("cris_mul_bcast", dg.ptr, dg.ld, s.ptr, s.ld, slot.ptr, slot.ld, f5a.rows, rpi, f5a.C)
            ds = self.f32(B * f5a.C)
            call("cris_mul_bcast_bwd_s", dg.ptr, dg.ld, f5a.ptr, f5a.ld, ds.data_ptr(), B, rpi, f5a.C)
            dsm = Mat(ds, B, f5a.C, fp32=True)
            self.accumulate_into(dsm, s)

        if self.training:
            self.on_backward(bwd_gate)
        f5 = self.bn_forward(g, n + ".norm_layer.0", True)
        fo2, fo1 = f5.C, self.P[n + ".f2_v_proj.0.weight"].shape[0]
        fo0 = self.P[n + ".f3_v_proj.0.weight"].shape[0]
        N4, H4, W4 = c4.geom
        cat2 = self.padded(N4, H4, W4, fo1 + fo2)
        self.conv_bn(c4, n + ".f2_v_proj.0.weight", n + ".f2_v_proj.1", 3, out=cat2.cols(0, fo1))
        self.upsample(f5, out=cat2.cols(fo1, fo1 + fo2))
        cat3 = self.padded(N4, H4, W4, fo0 + fo1)
        f4 = self.conv_bn(cat2, n + ".f2_cat.0.weight", n + ".f2_cat.1", 1, out=cat3.cols(fo0, fo0 + fo1))
        f3a = self.conv_bn(c3, n + ".f3_v_proj.0.weight", n + ".f3_v_proj.1", 3)
        self.avgpool(f3a, out=cat3.cols(0, fo0))
        f3 = self.conv_bn(cat3, n + ".f3_cat.0.weight", n + ".f3_cat.1", 1)
        cat4 = self.padded(N4, H4, W4, 3 * fo1)
        fq5 = self.conv_bn(f5, n + ".f4_proj5.0.weight", n + ".f4_proj5.1", 3)
        self.conv_bn(f4, n + ".f4_proj4.0.weight", n + ".f4_proj4.1", 3, out=cat4.cols(fo1, 2 * fo1))
        self.conv_bn(f3, n + ".f4_proj3.0.weight", n + ".f4_proj3.1", 3, out=cat4.cols(0, fo1))
        self.upsample(fq5, out=cat4.cols(2 * fo1, 3 * fo1))
        # CoordConv input: [fq | x | y | zero pad] (model/layers.py:30-39)
        cpad = _r8(fo1 + 2)
        cbuf = self.padded(N4, H4, W4, cpad)
        call("cris_coord_fill", cbuf.ptr, cbuf.ld, fo1, N4, H4, W4)
        self.conv_bn(cat4, n + ".aggr.0.weight", n + ".aggr.1", 1, out=cbuf.cols(0, fo1))
        cin_view = cbuf.cols(0, fo1 + 2)
        cin_view.need_grad = True
        fq = self.conv_bn(cin_view, n + ".coordconv.0.conv1.0.weight", n + ".coordconv.0.conv1.1", 3, cin=fo1 + 2)
        fq = self.conv_bn(fq, n + ".coordconv.1.0.weight", n + ".coordconv.1.1", 3)
        self.tap("fq", fq)
        return fq

    # ---- vision-language decoder: model/layers.py:154-250 ---------------------------------------------------------
    def decoder(self, fq: Mat, wfeat: Mat) -> Mat:
        B, H, W = fq.geom
        C = fq.C
        T = H * W
        L = self.word.shape[1]
        heads = self.model.num_head
        vpos = self.e.const(("pos2d", C, H, W), lambda: _pos2d(C, H, W), self.dev)
        tpos = self.e.const(("pos1d", C, L, B), lambda: _pos1d(C, L).repeat(B, 1), self.dev)
        vis = self.new(B * T, C, True)
        call("cris_padded_to_tokens", fq.ptr, fq.ld, None, 0, vis.ptr, 1, vis.ld, B, H, W, C)
        vis_in = vis  # `vis` is rebound by every decoder layer; the closure needs the first one

        def bwd_vis():
            g = self.grad_of(vis_in)
            if g is None:
                return
            slot, acc = self.grad_slot(fq)
            assert not acc
            call("cris_tokens_to_padded", g.ptr, 1, g.ld, slot.ptr, slot.ld, B, H, W, C)

        if self.training:
            self.on_backward(bwd_vis)
        # key input of the cross attention: txt + txt_pos (same for every layer)
        tk = self.new(B * L, C)
        tposm = Mat(tpos, B * L, C, fp32=True)
        self.ew(0, wfeat, tposm, tk)

        def bwd_tk():
            g = self.grad_of(tk)
            if g is not None:
                self.accumulate_into(g, wfeat)

        if self.training:
            self.on_backward(bwd_tk)
        p_drop = self.p_drop
        for i in range(len(self.model.decoder.layers)):
            p = f"decoder.layers.{i}"
            # self attention: q = k = LN(vis) + pos, v = LN(vis)
            v2, v2p = self.layernorm(vis, p + ".norm1", add=vpos, want_y=True, want_y2=True)
            qk = self.linear(v2p, p + ".self_attn.in_proj_weight", p + ".self_attn.in_proj_bias", w_rows=(0, 2 * C))
            vv = self.linear(v2, p + ".self_attn.in_proj_weight", p + ".self_attn.in_proj_bias", w_rows=(2 * C, 3 * C))
            a = self.attention(qk.cols(0, C), qk.cols(C, 2 * C), vv, B, heads, T, T, p_drop=p_drop)
            a = self.linear(a, p + ".self_attn.out_proj.weight", p + ".self_attn.out_proj.bias")
            a, _ = self.layernorm(a, p + ".self_attn_norm")
            vis = self.residual_add(vis, a, p_drop)
            # cross attention: q = LN(vis) + pos, k = txt + txt_pos, v = txt; padded tokens masked
            _, v2p = self.layernorm(vis, p + ".norm2", add=vpos, want_y=False, want_y2=True)
            q = self.linear(v2p, p + ".multihead_attn.in_proj_weight", p + ".multihead_attn.in_proj_bias", w_rows=(0, C))
            k = self.linear(tk, p + ".multihead_attn.in_proj_weight", p + ".multihead_attn.in_proj_bias",
                            w_rows=(C, 2 * C))
            vv = self.linear(wfeat, p + ".multihead_attn.in_proj_weight", p + ".multihead_attn.in_proj_bias",
                             w_rows=(2 * C, 3 * C))
            a = self.attention(q, k, vv, B, heads, T, L, key_pad=True, p_drop=p_drop)
            a = self.linear(a, p + ".multihead_attn.out_proj.weight", p + ".multihead_attn.out_proj.bias")
            a, _ = self.layernorm(a, p + ".cross_attn_norm")
            vis = self.residual_add(vis, a, p_drop)
            # FFN: Linear -> ReLU -> Dropout -> LayerNorm(ffn) -> Linear
            v2, _ = self.layernorm(vis, p + ".norm3")
            h = self.linear(v2, p + ".ffn.0.weight", p + ".ffn.0.bias", act=ACT_RELU)
            if p_drop > 0:
                h = self.dropout(h, p_drop)
            h, _ = self.layernorm(h, p + ".ffn.3")
            h = self.linear(h, p + ".ffn.4.weight", p + ".ffn.4.bias")
            vis = self.residual_add(vis, h, p_drop)
            self.tap(f"dec{i}", vis)
        out, _ = self.layernorm(vis, "decoder.norm")
        fo = self.padded(B, H, W, C)
        call("cris_tokens_to_padded", out.ptr, 0, out.ld, fo.ptr, fo.ld, B, H, W, C)

        def bwd_fo():
            g = self.grad_of(fo)
            if g is None:
                return
            slot, acc = self.grad_slot(out)
            assert not acc
            call("cris_padded_to_tokens", g.ptr, g.ld, None, 0, slot.ptr, 0, slot.ld, B, H, W, C)

        if self.training:
            self.on_backward(bwd_fo)
        self.tap("dec_out", fo)
        return fo

    def dropout(self, x: Mat, p: float) -> Mat:
        y = self.new(x.rows, x.C, False, None, ld=x.ld)
        sd = self.seed()
        self.ew(2, x, None, y, p, sd)

        def bwd():
            g = self.grad_of(y)
            if g is None:
                return
            slot, acc = self.grad_slot(x)
            assert not acc
            self.ew(2, g, None, slot, p, sd)

        if self.training:
            self.on_backward(bwd)
        return y

    # ---- projector + loss: model/layers.py:63-84, model/segmenter.py:52-62 ---------------------------------------
    def projector_and_loss(self, fq: Mat, state: Mat):
        B, H, W = fq.geom
        x = self.upsample(fq)
        x = self.conv_bn(x, "proj.vis.1.0.weight", "proj.vis.1.1", 3)
        x = self.upsample(x)
        x = self.conv_bn(x, "proj.vis.3.0.weight", "proj.vis.3.1", 3)
        xf, _, _ = self.conv(x, "proj.vis.4.weight", 1, stats=False, bias_name="proj.vis.4.bias")
        self.tap("proj_feat", xf)
        C = xf.C
        Ho, Wo = xf.geom[1], xf.geom[2]
        t = self.linear(state, "proj.txt.weight", "proj.txt.bias", out_fp32=True)
        pred = torch.empty(B, 1, Ho, Wo, dtype=torch.float32, device=self.dev)
        mask_out = torch.empty(B, 1, Ho, Wo, dtype=torch.float32, device=self.dev) if self.mask is not None else None
        loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        Hm, Wm = (self.mask.shape[-2], self.mask.shape[-1]) if self.mask is not None else (0, 0)
        # per-sample intersection / union counts of (sigmoid(pred) >= 0.35) vs (target != 0): trainMetricGPU's reduction
        # (utils/misc.py:114-129), taken inside the loss kernel; read through CRIS.train_metric()
        counts = torch.zeros(B, 2, dtype=torch.int32, device=self.dev) if self.mask is not None else None
        call("cris_dynconv_bce_fwd", xf.ptr, xf.ld, t.ptr, t.ld, self.mask.data_ptr() if self.mask is not None else None,
             Hm, Wm, pred.data_ptr(), mask_out.data_ptr() if mask_out is not None else None, loss.data_ptr(),
             counts.data_ptr() if counts is not None else None, 0.35, B, Ho, Wo, C)
        self.pred, self.mask_out, self.loss, self.metric_counts = pred, mask_out, loss, counts
        self.e.last_metric_counts = counts
        self._head = (xf, t, B, Ho, Wo, C)

    # =================================================================================================
    def backward(self, dloss: torch.Tensor, names: List[str]):
        self.backward_range(dloss, 0, None)
        return self.collect_grads(names)

    def backward_range(self, dloss: Optional[torch.Tensor], i0: int, i1: Optional[int]):
        """Run the backward closures with reversed-tape positions [i0, i1); position 0 is preceded by the loss head.
        The whole backward is backward_range(dloss, 0, None); GraphedStep may capture it in several ranges so that
        gradients of the late layers reach DistributedDataParallel while the early layers are still running."""
        self.in_backward = True
        self.ensure_pgrad()
        self._leaf_main, self._leaf_hold, self._leaf_pending = None, [], 0
        self._leaf_join_every = max(1, int(os.environ.get("CRIS_B200_WGRAD_JOIN", "64")))
        if (self.dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
                and os.environ.get("CRIS_B200_WGRAD_STREAM", "1") == "1"):
            self._leaf_main = torch.cuda.current_stream(self.dev)
        if i0 == 0:
            xf, t, B, Ho, Wo, C = self._head
            g = dloss.detach().float().reshape(1).contiguous()
            dl = self.f32(B * Ho * Wo)
            dxf, _ = self.grad_slot(xf)
            dt, _ = self.grad_slot(t)
            dt.buf.zero_()
            call("cris_dynconv_bce_bwd", xf.ptr, xf.ld, t.ptr, t.ld, self.pred.data_ptr(), self.mask_out.data_ptr(),
                 g.data_ptr(), dl.data_ptr(), dxf.ptr, dxf.ld, dt.ptr, dt.ld, B, Ho, Wo, C)
            self._rtape = list(reversed(self.tape))
            self.tape = []  # closures <-> Run form reference cycles; drop them so buffers are freed promptly
        n = len(self._rtape)
        i1 = n if i1 is None else min(i1, n)
        tr = getattr(self, "text_range", None)
        fork_at = (n - self.image_end) if tr is not None else -1   # reversed position of the image encoder's last closure
        forked = None
        for i in range(i0, i1):
            if i == fork_at and torch.cuda.is_current_stream_capturing():
                # everything the two towers need (FPN / decoder gradients) has been issued: the text tower's backward goes
                # to the side stream as a parallel branch, the image encoder's backward continues on this one
                main = torch.cuda.current_stream(self.dev)
                side = self.e.side_stream(self.dev)
                side.wait_stream(main)
                # The branch reads gradients that were ALLOCATED on this stream (word features, sentence state).  The
                # caching allocator would hand their memory back to this stream the moment the branch's closures drop
                # them — while the branch is still a concurrent part of the graph — so the closures (and through them
                # every tensor the branch touches) stay referenced until the join below.
                branch_refs = []
                with torch.cuda.stream(side):
                    for j in range(n - tr[1], n - tr[0]):
                        self.bwd_pos = j
                        if self._rtape[j] is not None:
                            self._rtape[j]()
                            branch_refs.append(self._rtape[j])
                            self._rtape[j] = None
                forked = (main, side, branch_refs)
            self.bwd_pos = i
            if self._rtape[i] is not None:
                self._rtape[i]()
                self._rtape[i] = None
        self.join_leaves()
        self._leaf_main = None
        if forked is not None:
            forked[0].wait_stream(forked[1])
            forked[2].clear()
        if i1 >= n:
            self._rtape = []

    def collect_grads(self, names: List[str]):
        out = []
        for k in names:
            out.append(self.pgrad.get(k))
        missing = [k for k, gk in zip(names, out) if gk is None]
        if missing:
            raise RuntimeError(f"cris.pytorch_b200 backward produced no gradient for {missing[:5]}")
        return out
