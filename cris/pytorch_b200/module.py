"""`CRIS` — the drop-in nn.Module surface of the reference's `model.segmenter.CRIS`
(model/segmenter.py:10-62) and `build_segmenter` (model/__init__.py:32-49).

The module tree below is a set of PARAMETER CONTAINERS: ordinary torch.nn modules arranged so that
`state_dict()` has exactly the reference's 662 (r50) names / shapes / dtypes,
`SyncBatchNorm.convert_sync_batchnorm`, `DistributedDataParallel`, `DataParallel`, `.cuda()`,
`.train()/.eval()` and `print(model)` behave as they do for the reference.  None of the containers'
own forwards is ever run: `CRIS.forward` hands the parameters to the sm_100a engine
(cris/pytorch_b200/engine.py -> libcris_b200.so).  There is no CPU / eager fallback.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn as nn


def _no_forward(*_a, **_k):
    raise RuntimeError("cris.pytorch_b200 containers hold parameters only; call CRIS.forward")


class _Holder(nn.Module):
    """A named bag of sub-modules / parameters (never called)."""
    forward = _no_forward


def _conv_bn_relu(cin: int, cout: int, k: int) -> nn.Sequential:
    # reference: conv_layer (model/layers.py:8-11) -> keys "0.weight", "1.{weight,bias,running_*}"
    return nn.Sequential(nn.Conv2d(cin, cout, k, 1, k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(True))


def _bottleneck(inplanes: int, planes: int, stride: int) -> _Holder:
    # reference: Bottleneck.__init__ (model/clip.py:13-42)
    m = _Holder()
    m.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
    m.bn1 = nn.BatchNorm2d(planes)
    m.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
    m.bn2 = nn.BatchNorm2d(planes)
    m.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
    m.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
    m.bn3 = nn.BatchNorm2d(planes * 4)
    m.relu = nn.ReLU(inplace=True)
    m.stride = stride
    m.downsample = None
    if stride > 1 or inplanes != planes * 4:
        m.downsample = nn.Sequential(OrderedDict([
            ("-1", nn.AvgPool2d(stride)),
            ("0", nn.Conv2d(inplanes, planes * 4, 1, stride=1, bias=False)),
            ("1", nn.BatchNorm2d(planes * 4))]))
    return m


def _visual(layers, output_dim: int, heads: int, spacial: int, width: int) -> _Holder:
    # reference: ModifiedResNet.__init__ + AttentionPool2d.__init__ (model/clip.py:61-78,154-196)
    v = _Holder()
    v.conv1 = nn.Conv2d(3, width // 2, 3, stride=2, padding=1, bias=False)
    v.bn1 = nn.BatchNorm2d(width // 2)
    v.conv2 = nn.Conv2d(width // 2, width // 2, 3, padding=1, bias=False)
    v.bn2 = nn.BatchNorm2d(width // 2)
    v.conv3 = nn.Conv2d(width // 2, width, 3, padding=1, bias=False)
    v.bn3 = nn.BatchNorm2d(width)
    v.avgpool = nn.AvgPool2d(2)
    v.relu = nn.ReLU(inplace=True)
    inpl = width
    for li, nb in enumerate(layers, start=1):
        planes = width * 2 ** (li - 1)
        blocks = []
        for bi in range(nb):
            blocks.append(_bottleneck(inpl, planes, 2 if (li > 1 and bi == 0) else 1))
            inpl = planes * 4
        setattr(v, f"layer{li}", nn.Sequential(*blocks))
    e = width * 32
    ap = _Holder()
    ap.positional_embedding = nn.Parameter(torch.randn(spacial ** 2 + 1, e) / e ** 0.5)
    ap.k_proj = nn.Linear(e, e)
    ap.q_proj = nn.Linear(e, e)
    ap.v_proj = nn.Linear(e, e)
    ap.c_proj = nn.Linear(e, output_dim)
    ap.num_heads = heads
    ap.spacial_dim = spacial
    ap.connect = nn.Sequential(nn.Conv2d(e, output_dim, 1, stride=1, bias=False), nn.BatchNorm2d(output_dim))
    v.attnpool = ap
    v.output_dim = output_dim
    return v


def _text_block(width: int, heads: int) -> _Holder:
    # reference: ResidualAttentionBlock.__init__ (model/clip.py:240-253)
    b = _Holder()
    b.attn = nn.MultiheadAttention(width, heads)
    b.ln_1 = nn.LayerNorm(width)
    b.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(width, width * 4)), ("gelu", nn.Identity()),
                                       ("c_proj", nn.Linear(width * 4, width))]))
    b.ln_2 = nn.LayerNorm(width)
    return b


class CLIPParams(_Holder):
    """Parameter tree of the reference's `CLIP` for the ResNet towers (model/clip.py:335-388)."""

    def __init__(self, embed_dim, spacial, vision_layers, vision_width, context_length, vocab_size, twidth, theads,
                 tlayers):
        super().__init__()
        self.context_length = context_length
        self.visual = _visual(vision_layers, embed_dim, vision_width * 32 // 64, spacial, vision_width)
        tr = _Holder()
        tr.width, tr.layers = twidth, tlayers
        tr.resblocks = nn.Sequential(*[_text_block(twidth, theads) for _ in range(tlayers)])
        self.transformer = tr
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, twidth)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, twidth).normal_(std=0.01))
        self.ln_final = nn.LayerNorm(twidth)
        self.text_projection = nn.Parameter(torch.empty(twidth, embed_dim).normal_(std=twidth ** -0.5))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))


# tensors the reference rounds through fp16 at import (convert_weights, model/clip.py:477-500)
def _fp16_rounded_at_import(key: str, model: nn.Module) -> bool:
    if key == "text_projection":
        return True
    mod_name, _, leaf = key.rpartition(".")
    try:
        mod = model.get_submodule(mod_name) if mod_name else model
    except AttributeError:
        return False
    if isinstance(mod, (nn.Conv2d, nn.Linear)) and leaf in ("weight", "bias"):
        return True
    if isinstance(mod, nn.MultiheadAttention) and leaf in ("in_proj_weight", "in_proj_bias"):
        return True
    return False


def build_clip_from_state_dict(sd: Dict[str, torch.Tensor], word_len: int) -> CLIPParams:
    """Same contract as the reference's `build_model(state_dict, txt_length)` (model/clip.py:503-554):
    architecture inferred from tensor shapes, conv/linear/MHA/text_projection values rounded through
    fp16, load with strict=False (the CRIS-added attnpool.connect.* keeps its fresh init), fp32 result."""
    if "visual.proj" in sd:
        raise NotImplementedError("ViT CLIP towers are not selected by any reference config (SURVEY.md §2 #2)")
    counts = []
    for b in (1, 2, 3, 4):
        counts.append(len({k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}.")}))
    vision_width = sd["visual.layer1.0.conv1.weight"].shape[0]
    n_pos = sd["visual.attnpool.positional_embedding"].shape[0]
    spacial = int(round((n_pos - 1) ** 0.5))
    if spacial * spacial + 1 != n_pos:
        raise ValueError("attnpool.positional_embedding is not (s*s+1) rows")
    embed_dim = sd["text_projection"].shape[1]
    twidth = sd["ln_final.weight"].shape[0]
    tlayers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")})
    clip = CLIPParams(embed_dim, spacial, tuple(counts), vision_width, sd["positional_embedding"].shape[0],
                      sd["token_embedding.weight"].shape[0], twidth, twidth // 64, tlayers)
    clip.word_len = word_len
    load = {}
    for k, v in sd.items():
        if k in ("input_resolution", "context_length", "vocab_size"):
            continue
        v = v.detach()
        if v.dtype.is_floating_point:
            v = v.float()
            if _fp16_rounded_at_import(k, clip):
                v = v.half().float()
        load[k] = v
    clip.load_state_dict(load, strict=False)
    return clip.float()


def _decoder_layer(d: int, heads: int, ffn: int, dropout: float) -> _Holder:
    # reference: TransformerDecoderLayer.__init__ (model/layers.py:192-219)
    m = _Holder()
    m.self_attn_norm = nn.LayerNorm(d)
    m.cross_attn_norm = nn.LayerNorm(d)
    m.self_attn = nn.MultiheadAttention(d, heads, dropout=dropout)
    m.multihead_attn = nn.MultiheadAttention(d, heads, dropout=dropout, kdim=d, vdim=d)
    m.ffn = nn.Sequential(nn.Linear(d, ffn), nn.ReLU(True), nn.Dropout(dropout), nn.LayerNorm(ffn), nn.Linear(ffn, d))
    m.norm1 = nn.LayerNorm(d)
    m.norm2 = nn.LayerNorm(d)
    m.norm3 = nn.LayerNorm(d)
    m.dropout1 = nn.Dropout(dropout)
    m.dropout2 = nn.Dropout(dropout)
    m.dropout3 = nn.Dropout(dropout)
    return m


class CRIS(nn.Module):
    """Drop-in for `model.segmenter.CRIS` (same ctor argument, forward signature, outputs, state_dict)."""

    def __init__(self, cfg):
        super().__init__()
        clip_sd = torch.jit.load(cfg.clip_pretrain, map_location="cpu").eval().state_dict()
        self.backbone = build_clip_from_state_dict(clip_sd, cfg.word_len)
        fi, fo = list(cfg.fpn_in), list(cfg.fpn_out)
        # reference: FPN.__init__ (model/layers.py:254-280)
        neck = _Holder()
        neck.txt_proj = nn.Sequential(nn.Linear(fi[2], fo[2], bias=False), nn.BatchNorm1d(fo[2]), nn.ReLU(True))
        neck.f1_v_proj = _conv_bn_relu(fi[2], fo[2], 1)
        neck.norm_layer = nn.Sequential(nn.BatchNorm2d(fo[2]), nn.ReLU(True))
        neck.f2_v_proj = _conv_bn_relu(fi[1], fo[1], 3)
        neck.f2_cat = _conv_bn_relu(fo[2] + fo[1], fo[1], 1)
        neck.f3_v_proj = _conv_bn_relu(fi[0], fo[0], 3)
        neck.f3_cat = _conv_bn_relu(fo[0] + fo[1], fo[1], 1)
        neck.f4_proj5 = _conv_bn_relu(fo[2], fo[1], 3)
        neck.f4_proj4 = _conv_bn_relu(fo[1], fo[1], 3)
        neck.f4_proj3 = _conv_bn_relu(fo[1], fo[1], 3)
        neck.aggr = _conv_bn_relu(3 * fo[1], fo[1], 1)
        coord = _Holder()
        coord.conv1 = _conv_bn_relu(fo[1] + 2, fo[1], 3)
        neck.coordconv = nn.Sequential(coord, _conv_bn_relu(fo[1], fo[1], 3))
        self.neck = neck
        # reference: TransformerDecoder.__init__ (model/layers.py:88-104)
        dec = _Holder()
        dec.layers = nn.ModuleList([_decoder_layer(cfg.vis_dim, cfg.num_head, cfg.dim_ffn, cfg.dropout)
                                    for _ in range(cfg.num_layers)])
        dec.num_layers = cfg.num_layers
        dec.norm = nn.LayerNorm(cfg.vis_dim)
        dec.return_intermediate = cfg.intermediate
        self.decoder = dec
        # reference: Projector.__init__ (model/layers.py:48-61)
        c = cfg.vis_dim // 2
        proj = _Holder()
        proj.in_dim, proj.kernel_size = c, 3
        proj.vis = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"), _conv_bn_relu(2 * c, 2 * c, 3),
                                 nn.Upsample(scale_factor=2, mode="bilinear"), _conv_bn_relu(2 * c, c, 3),
                                 nn.Conv2d(c, c, 1))
        proj.txt = nn.Linear(cfg.word_dim, c * 9 + 1)
        self.proj = proj
        if cfg.intermediate:
            raise NotImplementedError("cfg.intermediate=True is not used by any reference config")
        # the attention kernels are written for 64-wide heads (every reference config: 512/8, 2048/32, 512/8);
        # the reference derives head_dim = embed_dim // num_heads, so refuse anything else instead of mis-slicing
        if cfg.vis_dim % cfg.num_head != 0 or cfg.vis_dim // cfg.num_head != 64:
            raise ValueError(f"cris.pytorch_b200 needs vis_dim / num_head == 64 (got {cfg.vis_dim} / {cfg.num_head})")
        tw = self.backbone.ln_final.weight.shape[0]
        vw = self.backbone.visual.attnpool.q_proj.weight.shape[0]
        if tw % 64 != 0 or vw % 64 != 0 or vw // self.backbone.visual.attnpool.num_heads != 64:
            raise ValueError("cris.pytorch_b200 needs 64-wide attention heads in the CLIP text tower and attention pool")
        self.num_head = cfg.num_head
        self.dropout_p = float(cfg.dropout)
        self._engine = None

    # ------------------------------------------------------------------------------------------
    @property
    def _ddp_params_and_buffers_to_ignore(self):
        """Read by DistributedDataParallel at construction (torch/nn/parallel/distributed.py: names listed here are left
        out of its parameter reduction and of the per-forward buffer broadcast).  When EVERY BatchNorm of the tree is a
        SyncBatchNorm (train.py:97-98 converts before it wraps, :100-102) the running statistics are computed from the
        same rank-ordered global sums with the same arithmetic on every rank, i.e. they are bit-identical already, and
        DDP's default `broadcast_buffers=True` re-broadcast of all 213 buffers before every forward (coalesce + NCCL
        broadcast + 213 copy-back kernels, ~1.5 ms per step, profiles/r02_ddp_breakdown_n2.txt) is a no-op — so those
        buffers are listed.  With plain BatchNorm modules nothing is listed and DDP keeps rank 0's statistics
        authoritative exactly as it does for the reference."""
        bns = [m for m in self.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)]
        if not bns or not all(isinstance(m, nn.SyncBatchNorm) for m in bns):
            return []
        return [k for k, _ in self.named_buffers()
                if k.endswith(("running_mean", "running_var", "num_batches_tracked"))]

    def train_metric(self, pr_iou: float = 0.5):
        """(IoU, Pr@pr_iou) in percent of the LAST training forward — the numbers `trainMetricGPU(pred, target, 0.35,
        0.5)` (utils/misc.py:114-129, called at engine/engine.py:60) returns, from the per-sample intersection / union
        counts the loss kernel took while the logits were in registers (no extra pass over pred / target)."""
        c = self._get_engine().last_metric_counts
        if c is None:
            raise RuntimeError("train_metric() needs a preceding training-mode forward")
        c = c.float()
        ious = c[:, 0] / (c[:, 1] + 1e-6)
        return 100.0 * ious.mean(), 100.0 * (ious > pr_iou).float().mean()

    def _get_engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        return self._engine

    def forward(self, img, word, mask=None):
        """img [B,3,H,W] float, word [B,L] int64, mask [B,1,H,W] float (training).
        training: (pred.detach() [B,1,H/4,W/4], mask resized (nearest), loss); eval: pred.detach()
        — model/segmenter.py:29-62."""
        if not img.is_cuda:
            raise RuntimeError("cris.pytorch_b200.CRIS runs on a B200 GPU only (no CPU fallback)")
        return self._get_engine().run(img, word, mask)


def build_segmenter(args):
    """Reference: model/__init__.py:32-49 — (model, [backbone group, head group]) with `initial_lr`."""
    model = CRIS(args)
    backbone, head = [], []
    for k, v in model.named_parameters():
        if k.startswith("backbone") and "positional_embedding" not in k:
            backbone.append(v)
        else:
            head.append(v)
    try:
        from loguru import logger
        logger.info("Backbone with decay={}, Head={}".format(len(backbone), len(head)))
    except Exception:  # loguru is optional here
        pass
    param_list = [{"params": backbone, "initial_lr": args.lr_multi * args.base_lr},
                  {"params": head, "initial_lr": args.base_lr}]
    return model, param_list
