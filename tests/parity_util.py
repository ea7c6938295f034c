"""Shared helpers of the GPU parity tests (test infrastructure)."""
import os
import tempfile

import torch

from oracle import synth

THR = -0.6190392
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(arch, dropout=0.0):
    from cris.pytorch_b200 import CRIS
    cfg = synth.make_cfg(arch, dropout=dropout)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        cfg.clip_pretrain = path
        model = CRIS(cfg)
    sd = synth.full_state_dict(arch, 0, cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.cuda()


def group_of(name: str) -> str:
    p = name.split(".")
    if name.startswith("backbone.visual.layer"):
        return ".".join(p[:3])
    if name.startswith("backbone.visual.attnpool"):
        return "backbone.visual.attnpool"
    if name.startswith("backbone.visual"):
        return "backbone.visual.stem"
    if name.startswith("backbone.transformer.resblocks"):
        return "backbone.text." + ("0-5" if int(p[3]) < 6 else "6-11")
    if name.startswith("backbone"):
        return "backbone.text.embed"
    if name.startswith("decoder.layers"):
        return ".".join(p[:3])
    return p[0]


def gradient_report(named_grads, golden_grads, skip_prefixes=()):
    """named_grads: {name: fp32 gradient tensor (any device)}; golden: {name: {norm, idx, val} | None}.
    skip_prefixes: group / parameter-name prefixes left out of the per-block aggregates."""
    per, groups = {}, {}
    for k, gg in golden_grads.items():
        if gg is None:
            continue
        g = named_grads[k].detach().flatten()
        nref = float(gg["norm"])
        mine = g[gg["idx"].to(g.device)].double().cpu()
        ref = gg["val"].double()
        n = float(g.double().norm())
        per[k] = {"norm": n, "norm_ref": nref, "ratio": n / nref if nref > 1e-12 else None,
                  "sample_rel": float((mine - ref).norm() / (ref.norm() + 1e-30))}
        if nref < 1e-7:  # analytically zero in the reference (e.g. attnpool.k_proj.bias)
            continue
        if "txt_proj.1" in k:  # BatchNorm1d over the batch: a sign function at B=2, chaotic by construction
            continue
        if any(group_of(k).startswith(p) or k.startswith(p) for p in skip_prefixes):
            continue
        G = groups.setdefault(group_of(k), {"n2": 0.0, "r2": 0.0, "dot": 0.0, "m2": 0.0, "s2": 0.0, "count": 0})
        G["n2"] += n * n
        G["r2"] += nref * nref
        # samples are weighted by the tensor's size so that big tensors dominate like they do in the full gradient
        w = g.numel() / max(1, mine.numel())
        G["dot"] += w * float(mine @ ref)
        G["m2"] += w * float(mine @ mine)
        G["s2"] += w * float(ref @ ref)
        G["count"] += 1
    out = {}
    for name, G in groups.items():
        out[name] = {"norm_ratio": (G["n2"] / G["r2"]) ** 0.5, "cos": G["dot"] / ((G["m2"] * G["s2"]) ** 0.5 + 1e-30),
                     "tensors": G["count"]}
    return per, out
