"""Deterministic synthetic weights / inputs for CRIS (test + bench infrastructure).

There is no network: OpenAI's RN50.pt / RN101.pt and RefCOCO are unavailable, so every parity
test and benchmark runs on seeded synthetic weights of the exact reference shapes
 and seeded synthetic inputs of the dataset's shape.
CPU `torch.Generator` streams are bit-reproducible across machines for a fixed torch build,
which is what lets golden OUTPUTS be committed without committing 590 MB of weights.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict

import torch

ARCHS = {
    # name: (resnet blocks, embed_dim(word_dim), vision_width, text width, text layers, vocab, ctx)
    "r50": dict(layers=(3, 4, 6, 3), embed_dim=1024, width=64, twidth=512, tlayers=12, vocab=49408, ctx=77,
                spacial=7),
    "r101": dict(layers=(3, 4, 23, 3), embed_dim=512, width=64, twidth=512, tlayers=12, vocab=49408, ctx=77,
                 spacial=7),
    # reduced-width variant for fast CPU tests (same topology / code paths, ~1/16 the FLOPs)
    "tiny": dict(layers=(1, 2, 1, 1), embed_dim=128, width=16, twidth=128, tlayers=2, vocab=512, ctx=77,
                 spacial=7),
}


def make_cfg(arch: str = "r50", word_len: int = 17, dropout: float = 0.0, clip_pretrain: str = ""):
    """The TRAIN-section keys CRIS reads (config/refcoco/cris_r50.yaml:10-23)."""
    a = ARCHS[arch]
    w = a["width"]
    vis_dim = 8 * w  # 512 for width 64
    return SimpleNamespace(
        clip_pretrain=clip_pretrain, input_size=416, word_len=word_len, word_dim=a["embed_dim"], vis_dim=vis_dim,
        fpn_in=[8 * w, 16 * w, a["embed_dim"]], fpn_out=[4 * w, 8 * w, 16 * w], sync_bn=True,
        num_layers=3, num_head=8 if arch != "tiny" else 2, dim_ffn=32 * w, dropout=dropout, intermediate=False,
        lr_multi=0.1, base_lr=1e-4)


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def normal(self, *shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=self.g) * std + mean

    def uniform(self, *shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=self.g) * (hi - lo) + lo


def _bn(sd, g: _Gen, name: str, c: int, gamma=(0.6, 1.4)):
    sd[name + ".weight"] = g.uniform(c, lo=gamma[0], hi=gamma[1])
    sd[name + ".bias"] = g.normal(c, std=0.1)
    sd[name + ".running_mean"] = g.normal(c, std=0.1)
    sd[name + ".running_var"] = g.uniform(c, lo=0.6, hi=1.4)
    sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _conv(sd, g: _Gen, name: str, cout: int, cin: int, k: int, gain: float = 2.0):
    sd[name] = g.normal(cout, cin, k, k, std=math.sqrt(gain / (cin * k * k)))


def _linear(sd, g: _Gen, name: str, cout: int, cin: int, bias=True, std=None):
    sd[name + ".weight"] = g.normal(cout, cin, std=std if std is not None else 1.0 / math.sqrt(cin))
    if bias:
        sd[name + ".bias"] = g.normal(cout, std=0.02)


def _ln(sd, g: _Gen, name: str, c: int):
    sd[name + ".weight"] = g.uniform(c, lo=0.8, hi=1.2)
    sd[name + ".bias"] = g.normal(c, std=0.05)


def clip_state_dict(arch: str = "r50", seed: int = 0) -> Dict[str, torch.Tensor]:
    """A CLIP-ResNet-shaped state_dict with the key set OpenAI's TorchScript files carry PLUS the
    CRIS-added `visual.attnpool.connect.*` (model/clip.py:76-78); names per SURVEY.md Appendix B
    without the `backbone.` prefix.

    Conditioning: the last BatchNorm of every bottleneck gets a small gamma (0.05-0.15), as in trained /
    zero-init-residual ResNets (CLIP's own init zeroes it, model/clip.py:402-408).  With O(1) gammas a
    random-weight ResNet in batch-statistics mode is chaotic: bf16 storage noise grows ~2x per stage
    (12 % at layer4), for the reference under autocast as much as for this build, which would make
    any fp32-vs-bf16 comparison meaningless."""
    a = ARCHS[arch]
    g = _Gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    w = a["width"]
    # stem (model/clip.py:165-183)
    _conv(sd, g, "visual.conv1.weight", w // 2, 3, 3); _bn(sd, g, "visual.bn1", w // 2)
    _conv(sd, g, "visual.conv2.weight", w // 2, w // 2, 3); _bn(sd, g, "visual.bn2", w // 2)
    _conv(sd, g, "visual.conv3.weight", w, w // 2, 3); _bn(sd, g, "visual.bn3", w)
    inpl = w
    for li, nb in enumerate(a["layers"], start=1):
        planes = w * (2 ** (li - 1))
        for bi in range(nb):
            p = f"visual.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            _conv(sd, g, p + ".conv1.weight", planes, inpl, 1); _bn(sd, g, p + ".bn1", planes)
            _conv(sd, g, p + ".conv2.weight", planes, planes, 3); _bn(sd, g, p + ".bn2", planes)
            _conv(sd, g, p + ".conv3.weight", planes * 4, planes, 1); _bn(sd, g, p + ".bn3", planes * 4, gamma=(0.05, 0.15))
            if stride > 1 or inpl != planes * 4:
                _conv(sd, g, p + ".downsample.0.weight", planes * 4, inpl, 1, gain=1.0)
                _bn(sd, g, p + ".downsample.1", planes * 4)
            inpl = planes * 4
    e = w * 32
    ap = "visual.attnpool"
    sd[ap + ".positional_embedding"] = g.normal(a["spacial"] ** 2 + 1, e, std=e ** -0.5)
    for nm in ("q_proj", "k_proj", "v_proj"):
        _linear(sd, g, f"{ap}.{nm}", e, e, std=e ** -0.5)
    _linear(sd, g, ap + ".c_proj", a["embed_dim"], e, std=e ** -0.5)
    _conv(sd, g, ap + ".connect.0.weight", a["embed_dim"], e, 1, gain=1.0); _bn(sd, g, ap + ".connect.1", a["embed_dim"])
    # text tower (model/clip.py:371-388, init :410-422)
    tw, tl = a["twidth"], a["tlayers"]
    sd["token_embedding.weight"] = g.normal(a["vocab"], tw, std=0.02)
    sd["positional_embedding"] = g.normal(a["ctx"], tw, std=0.01)
    proj_std = (tw ** -0.5) * ((2 * tl) ** -0.5)
    for i in range(tl):
        p = f"transformer.resblocks.{i}"
        sd[p + ".attn.in_proj_weight"] = g.normal(3 * tw, tw, std=tw ** -0.5)
        sd[p + ".attn.in_proj_bias"] = g.normal(3 * tw, std=0.02)
        _linear(sd, g, p + ".attn.out_proj", tw, tw, std=proj_std)
        _ln(sd, g, p + ".ln_1", tw)
        _linear(sd, g, p + ".mlp.c_fc", 4 * tw, tw, std=(2 * tw) ** -0.5)
        _linear(sd, g, p + ".mlp.c_proj", tw, 4 * tw, std=proj_std)
        _ln(sd, g, p + ".ln_2", tw)
    _ln(sd, g, "ln_final", tw)
    sd["text_projection"] = g.normal(tw, a["embed_dim"], std=tw ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


def head_state_dict(cfg, seed: int = 1) -> Dict[str, torch.Tensor]:
    """neck / decoder / proj parameters (model/layers.py ctors) with seeded values."""
    g = _Gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    fi, fo = cfg.fpn_in, cfg.fpn_out

    def cl(name, cin, cout, k):
        _conv(sd, g, name + ".0.weight", cout, cin, k); _bn(sd, g, name + ".1", cout)

    sd["neck.txt_proj.0.weight"] = g.normal(fo[2], fi[2], std=1.0 / math.sqrt(fi[2]))
    _bn(sd, g, "neck.txt_proj.1", fo[2])
    cl("neck.f1_v_proj", fi[2], fo[2], 1)
    _bn(sd, g, "neck.norm_layer.0", fo[2])
    cl("neck.f2_v_proj", fi[1], fo[1], 3)
    cl("neck.f2_cat", fo[2] + fo[1], fo[1], 1)
    cl("neck.f3_v_proj", fi[0], fo[0], 3)
    cl("neck.f3_cat", fo[0] + fo[1], fo[1], 1)
    cl("neck.f4_proj5", fo[2], fo[1], 3)
    cl("neck.f4_proj4", fo[1], fo[1], 3)
    cl("neck.f4_proj3", fo[1], fo[1], 3)
    cl("neck.aggr", 3 * fo[1], fo[1], 1)
    cl("neck.coordconv.0.conv1", fo[1] + 2, fo[1], 3)
    cl("neck.coordconv.1", fo[1], fo[1], 3)
    d, ff = cfg.vis_dim, cfg.dim_ffn
    for i in range(cfg.num_layers):
        p = f"decoder.layers.{i}"
        for att in ("self_attn", "multihead_attn"):
            sd[f"{p}.{att}.in_proj_weight"] = g.normal(3 * d, d, std=d ** -0.5)
            sd[f"{p}.{att}.in_proj_bias"] = g.normal(3 * d, std=0.02)
            _linear(sd, g, f"{p}.{att}.out_proj", d, d)
        for nm in ("norm1", "norm2", "norm3", "self_attn_norm", "cross_attn_norm"):
            _ln(sd, g, f"{p}.{nm}", d)
        _linear(sd, g, p + ".ffn.0", ff, d, std=math.sqrt(2.0 / d))
        _ln(sd, g, p + ".ffn.3", ff)
        _linear(sd, g, p + ".ffn.4", d, ff)
    _ln(sd, g, "decoder.norm", d)
    c = cfg.vis_dim // 2
    cl("proj.vis.1", 2 * c, 2 * c, 3)
    cl("proj.vis.3", 2 * c, c, 3)
    _conv(sd, g, "proj.vis.4.weight", c, c, 1, gain=1.0)
    sd["proj.vis.4.bias"] = g.normal(c, std=0.02)
    # small dynamic kernels keep the logits O(1) and centred near the 0.35 threshold
    # r101 (word_dim 512): 4x the kernel scale, otherwise the logits' spread (sigma 0.07) sits inside the +-0.05 band
    gain = 4.0 if (cfg.vis_dim, cfg.word_dim) == (512, 512) else 1.0
    sd["proj.txt.weight"] = g.normal(c * 9 + 1, cfg.word_dim, std=gain * 0.35 / math.sqrt(cfg.word_dim * c * 9))
    sd["proj.txt.bias"] = g.normal(c * 9 + 1, std=0.002)
    # centre the eval logits on the mask threshold sigmoid(l) > 0.35 <=> l > -0.619 so that thresholded-mask /
    # IoU comparisons are not vacuous; offsets measured once per arch with these seeds
    sd["proj.txt.bias"][-1] = {(512, 1024): 0.10, (512, 512): -0.435, (128, 128): -0.52}.get((cfg.vis_dim, cfg.word_dim), 0.10)
    return sd


def full_state_dict(arch: str = "r50", seed: int = 0, cfg=None) -> Dict[str, torch.Tensor]:
    """All 662 (r50) entries of CRIS.state_dict(), reference names."""
    cfg = cfg or make_cfg(arch)
    sd = {"backbone." + k: v for k, v in clip_state_dict(arch, seed).items()}
    sd.update(head_state_dict(cfg, seed + 1))
    return sd


def make_inputs(batch: int, seed: int = 0, size: int = 416, word_len: int = 17, vocab: int = 49408):
    """Synthetic RefCOCO-shaped batch: CLIP-normalised image ~N(0,1); tokens
    SOT + n content tokens + EOT + zero padding (utils/dataset.py:67-82); soft-edged mask in [0,1]."""
    g = torch.Generator().manual_seed(1000 + seed)
    img = torch.randn(batch, 3, size, size, generator=g)
    sot, eot = vocab - 2, vocab - 1
    word = torch.zeros(batch, word_len, dtype=torch.long)
    mask = torch.zeros(batch, 1, size, size)
    yy = torch.arange(size, dtype=torch.float32)[:, None]
    xx = torch.arange(size, dtype=torch.float32)[None, :]
    for b in range(batch):
        n = int(torch.randint(3, word_len - 1, (1,), generator=g))
        word[b, 0] = sot
        word[b, 1:n + 1] = torch.randint(1, vocab - 2, (n,), generator=g)
        word[b, n + 1] = eot
        cy, cx = (torch.rand(2, generator=g) * 0.6 + 0.2) * size
        ry, rx = (torch.rand(2, generator=g) * 0.25 + 0.08) * size
        d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        mask[b, 0] = torch.clamp((1.15 - d) * 4.0, 0, 1)
    return img, word, mask


def save_clip_torchscript(sd: Dict[str, torch.Tensor], path: str, drop_connect: bool = True) -> None:
    """Write a TorchScript file whose `.state_dict()` equals `sd` — a stand-in for OpenAI's RN50.pt that
    `torch.jit.load(cfg.clip_pretrain).state_dict()` (model/segmenter.py:14-15) can consume.  OpenAI files do
    not carry the CRIS-added `visual.attnpool.connect.*`, so those keys are dropped by default."""
    root = torch.nn.Module()
    for k, v in sd.items():
        if drop_connect and ".connect." in k:
            continue
        parts = k.split(".")
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, torch.nn.Module())
            m = getattr(m, p)
        if v.dtype.is_floating_point and not parts[-1].startswith("running_"):
            m.register_parameter(parts[-1], torch.nn.Parameter(v.clone(), requires_grad=False))
        else:
            m.register_buffer(parts[-1], v.clone())
    torch.jit.save(torch.jit.script(root), path)
