"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only), writes tests/golden/

Imports `/root/reference/model` as-is, feeds it the seeded synthetic weights/inputs of
oracle/synth.py and records its outputs.  The fixtures pin (a) the oracle restatement
(tests/test_oracle_golden.py, CPU) and (b) the CUDA path (tests/test_parity_gpu.py) on a box
where /root/reference does not exist.  Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")


def scripted_param_tree(sd):
    """A TorchScript module whose state_dict() is exactly `sd` (what torch.jit.load(RN50.pt)
    returns for the real checkpoint) — model/segmenter.py:14-15 only calls .state_dict() on it."""
    root = torch.nn.Module()
    for k, v in sd.items():
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
    return torch.jit.script(root)


def import_reference():
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    import model as ref_model  # noqa
    from model.segmenter import CRIS as RefCRIS
    sys.path.remove(REF)
    return RefCRIS


def sample_indices(numel: int, n: int = 24):
    g = torch.Generator().manual_seed(numel % 100003)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def run(arch: str, batch: int, size: int, tag: str, RefCRIS):
    t0 = time.time()
    cfg = synth.make_cfg(arch)
    clip_sd = synth.clip_state_dict(arch, seed=0)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        # OpenAI checkpoints do not carry the CRIS-added connect.* keys
        torch.jit.save(scripted_param_tree({k: v for k, v in clip_sd.items() if ".connect." not in k}), path)
        cfg.clip_pretrain = path
        torch.manual_seed(0)
        ref = RefCRIS(cfg)
    # constructor contract: fp16 rounding of conv/linear/MHA/text_projection weights (clip.py:477-500)
    ctor_sd = ref.state_dict()
    ctor = {"keys": list(ctor_sd.keys()), "shapes": {k: tuple(v.shape) for k, v in ctor_sd.items()},
            "dtypes": {k: str(v.dtype) for k, v in ctor_sd.items()},
            "fp16_rounded": {k: bool(torch.equal(v, clip_sd[k[9:]].half().float()))
                             for k, v in ctor_sd.items() if k.startswith("backbone.") and k[9:] in clip_sd
                             and v.dtype.is_floating_point and ".connect." not in k},
            "unchanged": {k: bool(torch.equal(v, clip_sd[k[9:]]))
                          for k, v in ctor_sd.items() if k.startswith("backbone.") and k[9:] in clip_sd
                          and ".connect." not in k}}
    full = synth.full_state_dict(arch, seed=0, cfg=cfg)
    missing = ref.load_state_dict(full, strict=True)
    img, word, mask = synth.make_inputs(batch, seed=0, size=size, word_len=cfg.word_len,
                                        vocab=synth.ARCHS[arch]["vocab"])
    out = {"arch": arch, "batch": batch, "size": size, "ctor": ctor, "strict_load": str(missing)}
    ref.eval()
    with torch.no_grad():
        out["eval_pred"] = ref(img, word).clone()
    ref.train()
    pred, m, loss = ref(img, word, mask)
    loss.backward()
    out["train_pred"] = pred.clone()
    out["train_mask"] = m.clone()
    out["train_loss"] = loss.detach().clone()
    grads = {}
    for k, p in ref.named_parameters():
        if p.grad is None:
            grads[k] = None
            continue
        gflat = p.grad.flatten()
        idx = sample_indices(gflat.numel())
        grads[k] = {"norm": gflat.double().norm().float(), "idx": idx, "val": gflat[idx].clone()}
    out["grads"] = grads
    run_stats = {}
    for k, v in ref.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            run_stats[k] = {"norm": v.double().norm().float(), "head": v.flatten()[:4].clone()}
    out["running"] = run_stats
    os.makedirs(OUT, exist_ok=True)
    torch.save(out, os.path.join(OUT, f"{tag}.pt"))
    print(f"[golden] {tag}: eval_pred mean {out['eval_pred'].mean():.4f} std {out['eval_pred'].std():.4f} "
          f"train loss {float(loss):.6f}  ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    RefCRIS = import_reference()
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    cases = [("tiny", 2, 128, "tiny_b2_128"), ("r50", 2, 416, "r50_b2_416"),
             # round 2: the batch the SyncBN/DDP equivalence test shards over two ranks, and the r101 config
             ("r50", 8, 416, "r50_b8_416"), ("r101", 4, 416, "r101_b4_416")]
    for arch, b, size, tag in cases:
        if only and tag not in only:
            continue
        if "--tiny-only" in sys.argv and arch != "tiny":
            continue
        run(arch, b, size, tag, RefCRIS)
