"""CPU: the drop-in module surface — state_dict names/shapes/dtypes/order, checkpoint-import rounding, param groups —
against the constructor contract recorded from the unmodified reference (tests/golden/*.pt)."""
import os
import tempfile

import torch

from cris.pytorch_b200.module import build_segmenter
from oracle import synth


def _build(arch):
    cfg = synth.make_cfg(arch)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        cfg.clip_pretrain = path
        return cfg, build_segmenter(cfg)


def test_state_dict_matches_reference_constructor(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_b2_128.pt"), weights_only=False)["ctor"]
    cfg, (model, groups) = _build("tiny")
    sd = model.state_dict()
    assert list(sd.keys()) == g["keys"]
    for k, v in sd.items():
        assert tuple(v.shape) == g["shapes"][k], k
        assert str(v.dtype) == g["dtypes"][k], k
    clip = synth.clip_state_dict("tiny", 0)
    for k, rounded in g["fp16_rounded"].items():
        ref = clip[k[len("backbone."):]]
        exp = ref.half().float() if (rounded and not g["unchanged"][k]) else ref
        assert torch.equal(sd[k], exp), k
    # param groups: model/__init__.py:36-48
    names = dict(model.named_parameters())
    n_backbone = sum(1 for k in names if k.startswith("backbone") and "positional_embedding" not in k)
    assert len(groups[0]["params"]) == n_backbone and len(groups[1]["params"]) == len(names) - n_backbone
    assert groups[0]["initial_lr"] == cfg.lr_multi * cfg.base_lr and groups[1]["initial_lr"] == cfg.base_lr
    # strict round trip, and the `module.`-prefixed checkpoint layout of train.py:200 / test.py:78
    full = synth.full_state_dict("tiny", 0, cfg)
    model.load_state_dict(full, strict=True)
    wrapped = torch.nn.DataParallel(model)
    wrapped.load_state_dict({"module." + k: v for k, v in full.items()}, strict=True)


def test_syncbn_conversion_keeps_parameter_names():
    cfg, (model, _) = _build("tiny")
    keys = list(model.state_dict().keys())
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    assert list(conv.state_dict().keys()) == keys
    assert any(isinstance(m, torch.nn.SyncBatchNorm) for m in conv.modules())
    assert conv.training and conv.backbone.visual.bn1.weight.requires_grad
