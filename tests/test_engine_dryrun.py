"""CPU: dry-run of the engine's host logic (launch sequencing, shapes, gradient bookkeeping) with the kernel
launches replaced by argument-checking stubs — no compute happens here (that is what `-m gpu` tests check).
It verifies that one forward + backward of the full model issues only well-formed C-ABI calls and that every
parameter except `backbone.logit_scale` receives a gradient buffer of the right shape."""
import ctypes as C
import os
import tempfile

import pytest
import torch

from cris.pytorch_b200 import _lib, engine as eng_mod
from cris.pytorch_b200.module import CRIS
from oracle import synth


class _Recorder:
    def __init__(self):
        self.calls = []

    def call(self, name, *args):
        sig = _lib._SIGS[name]
        assert len(args) == len(sig) - 1, f"{name}: {len(args)} args for signature {sig}"
        for ch, a in zip(sig, args):
            if ch == "p":
                assert a is None or isinstance(a, int), (name, a)
            else:
                _lib._T[ch](a)  # ctypes conversion must succeed
        self.calls.append(name)

    def gemm(self, g):
        assert isinstance(g, _lib.GemmArgs)
        assert g.M > 0 and g.N > 0 and g.K > 0 and g.A and g.B and g.D
        assert g.lda % 8 == 0 and g.ldb % 8 == 0, (g.lda, g.ldb)
        if not g.d_fp32 and g.d_col_stride <= 1:
            assert g.ldd % 8 == 0
        if g.splits > 1:
            assert g.d_fp32 and g.accumulate
        self.calls.append("cris_gemm")


@pytest.mark.parametrize("training", [False, True])
def test_engine_host_logic(monkeypatch, training):
    rec = _Recorder()
    monkeypatch.setattr(eng_mod, "call", rec.call)
    monkeypatch.setattr(eng_mod, "gemm", rec.gemm)
    monkeypatch.setattr(_lib, "device_check", lambda: None)
    cfg = synth.make_cfg("tiny", dropout=0.1)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict("tiny", 0), path)
        cfg.clip_pretrain = path
        model = CRIS(cfg)
    model.train(training)
    img, word, mask = synth.make_inputs(2, 0, 128, cfg.word_len, synth.ARCHS["tiny"]["vocab"])
    e = model._get_engine()
    e.use_graphs = False  # CUDA-graph capture needs a GPU; the eager path issues the same launches
    if not training:
        pred = e.run(img, word, None)
        assert pred.shape == (2, 1, 32, 32)
        assert rec.calls.count("cris_gemm") > 50
        return
    pred, m, loss = e.run(img, word, mask)
    assert pred.shape == (2, 1, 32, 32) and m.shape == (2, 1, 32, 32) and loss.dim() == 0
    n_fwd = len(rec.calls)
    loss.backward()
    assert len(rec.calls) > 2 * n_fwd * 0.8
    for k, p in model.named_parameters():
        if k == "backbone.logit_scale":
            assert p.grad is None
        else:
            assert p.grad is not None and p.grad.shape == p.shape, k
