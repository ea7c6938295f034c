"""Checkpoint-format tooling.  The reference writes `last_model.pth` / `best_model.pth` as
{'epoch', 'cur_iou', 'best_iou', 'prec', 'state_dict' (keys prefixed `module.` by DistributedDataParallel), 'optimizer',
'scheduler'} (train.py:192-207), resumes from it (train.py:160-176) and tests from `best_model.pth` through
`torch.nn.DataParallel(model).load_state_dict(checkpoint['state_dict'], strict=True)` (test.py:71-79).

  python tools/checkpoint_tool.py inspect  ckpt.pth                  # keys, shapes, prefix, optimizer layout
  python tools/checkpoint_tool.py verify   ckpt.pth --arch r50       # strict-load check against cris.pytorch_b200.CRIS
  python tools/checkpoint_tool.py strip    in.pth out.pth            # drop the `module.` prefix (bare-module loading)
  python tools/checkpoint_tool.py wrap     in.pth out.pth            # add it (so test.py's DataParallel load works)
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CKPT_KEYS = ("epoch", "cur_iou", "best_iou", "prec", "state_dict", "optimizer", "scheduler")


def save_checkpoint(path, model, optimizer, scheduler, epoch, cur_iou, best_iou, prec):
    """Write exactly the dictionary of train.py:192-207 (model may be the DDP / DataParallel wrapper)."""
    torch.save({"epoch": epoch, "cur_iou": cur_iou, "best_iou": best_iou, "prec": prec, "state_dict": model.state_dict(),
                "optimizer": optimizer.state_dict(), "scheduler": scheduler.state_dict()}, path)


def strip_prefix(sd, prefix="module."):
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}


def add_prefix(sd, prefix="module."):
    return {(k if k.startswith(prefix) else prefix + k): v for k, v in sd.items()}


def inspect(ck):
    sd = ck["state_dict"] if "state_dict" in ck else ck
    pref = all(k.startswith("module.") for k in sd)
    n_param = sum(v.numel() for v in sd.values() if v.dtype.is_floating_point)
    info = {"checkpoint_keys": [k for k in CKPT_KEYS if k in ck], "entries": len(sd), "module_prefix": pref,
            "float_elements_M": round(n_param / 1e6, 2)}
    if "optimizer" in ck:
        o = ck["optimizer"]
        info["optimizer_groups"] = [len(g["params"]) for g in o["param_groups"]]
        info["optimizer_state_entries"] = len(o["state"])
    return info


def verify(ck, arch):
    from oracle import synth
    import tempfile
    from cris.pytorch_b200 import build_segmenter
    cfg = synth.make_cfg(arch)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), p)
        cfg.clip_pretrain = p
        model, _ = build_segmenter(cfg)
    sd = ck["state_dict"] if "state_dict" in ck else ck
    torch.nn.DataParallel(model).load_state_dict(add_prefix(sd), strict=True)   # test.py:71-78
    model.load_state_dict(strip_prefix(sd), strict=True)
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["inspect", "verify", "strip", "wrap"])
    ap.add_argument("src")
    ap.add_argument("dst", nargs="?")
    ap.add_argument("--arch", default="r50")
    a = ap.parse_args()
    ck = torch.load(a.src, map_location="cpu", weights_only=False)
    if a.cmd == "inspect":
        print(inspect(ck))
    elif a.cmd == "verify":
        print("strict load ok" if verify(ck, a.arch) else "FAILED")
    else:
        fn = strip_prefix if a.cmd == "strip" else add_prefix
        if "state_dict" in ck:
            ck["state_dict"] = fn(ck["state_dict"])
        else:
            ck = fn(ck)
        torch.save(ck, a.dst)
        print(f"wrote {a.dst}")


if __name__ == "__main__":
    main()
