"""Drive the UNMODIFIED reference (staged at baseline/_ref by tools/stage_reference.py) through its own public API.

Measurement infrastructure only (bench.py's `--impl reference`, `cpu_baseline` and `gpu_incumbent` legs, and the
drop-in tests): it imports `model.build_segmenter` from baseline/_ref exactly as the reference's train.py does,
feeds it the seeded synthetic CLIP TorchScript file + synthetic batch of oracle/synth.py, and runs the training
iteration of the reference's engine/engine.py:48-70 (autocast forward, zero_grad, scaled backward, scaler.step,
scaler.update, trainMetricGPU, three host reads).  Nothing of cris.pytorch_b200 is on this path.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model", "segmenter.py"))


def import_reference():
    """-> (build_segmenter, trainMetricGPU, load_cfg_from_cfg_file) of the staged reference."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python tools/stage_reference.py` in the build container")
    for k in [k for k in sys.modules if k in ("model", "utils") or k.startswith(("model.", "utils."))]:
        m = sys.modules[k]
        if not getattr(m, "__file__", "") or not str(getattr(m, "__file__", "")).startswith(REF):
            del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        from model import build_segmenter
        from utils.config import load_cfg_from_cfg_file
        from utils.misc import trainMetricGPU
    finally:
        sys.path.remove(REF)
    return build_segmenter, trainMetricGPU, load_cfg_from_cfg_file


def load_cfg(arch: str, dropout=None):
    """The reference's own yaml (config/refcoco/cris_<arch>.yaml) through its own loader; 'tiny' = the synthetic
    reduced-width config of oracle/synth.py (no yaml exists for it)."""
    from oracle import synth
    if arch == "tiny":
        cfg = synth.make_cfg("tiny", dropout=0.1 if dropout is None else dropout)
        cfg.max_norm, cfg.weight_decay = 0.0, 0.0
        return cfg
    _, _, load = import_reference()
    cfg = load(os.path.join(REF, "config", "refcoco", f"cris_{arch}.yaml"))
    if dropout is not None:
        cfg.dropout = dropout
    return cfg


def build_reference_model(arch: str, dropout=None):
    """(cfg, model, param groups): the reference's CRIS with the synthetic weights of oracle/synth.py."""
    from oracle import synth
    build_segmenter, _, _ = import_reference()
    cfg = load_cfg(arch, dropout)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        cfg.clip_pretrain = path
        model, groups = build_segmenter(cfg)
    model.load_state_dict(synth.full_state_dict(arch, 0, synth.make_cfg(arch)), strict=True)
    return cfg, model, groups


def train_steps(arch: str, batch: int, steps: int, warmup: int, device: str = "cpu", amp_dtype=None, threads=None,
                size=None):
    """Time `steps` reference training iterations (engine/engine.py:48-70).  -> dict(seconds, batch, losses).
    device 'cpu': fp32 (torch.cuda.amp.autocast is a no-op without CUDA, as in the reference on a CPU box).
    device 'cuda': autocast(amp_dtype) + GradScaler (enabled for fp16 = the stock recipe) + cudnn.benchmark."""
    from oracle import synth
    _, metric, _ = import_reference()
    if threads:
        torch.set_num_threads(threads)
    cfg, model, groups = build_reference_model(arch)
    size = size or (128 if arch == "tiny" else 416)
    dev = torch.device(device)
    on_gpu = dev.type == "cuda"
    if on_gpu:
        torch.backends.cudnn.benchmark = True  # tools/latency.py:46, train.py's cudnn defaults
    model = model.to(dev).train()
    opt = torch.optim.Adam(groups, lr=cfg.base_lr, weight_decay=cfg.weight_decay)  # train.py:105-107
    use_scaler = on_gpu and amp_dtype == torch.float16
    scaler = torch.amp.GradScaler("cuda", enabled=use_scaler)                       # train.py:111
    img, word, mask = synth.make_inputs(batch, 0, size, cfg.word_len, synth.ARCHS[arch]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)
    losses = []

    def step():
        with torch.autocast(dev.type, dtype=amp_dtype, enabled=on_gpu and amp_dtype is not None):
            pred, target, loss = model(img, word, mask)             # engine.py:48-49
        opt.zero_grad()                                             # :52
        scaler.scale(loss).backward()                               # :53
        if getattr(cfg, "max_norm", 0):
            torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.max_norm)
        scaler.step(opt)                                            # :56
        scaler.update()                                             # :57
        iou, pr5 = metric(pred.float(), target, 0.35, 0.5)          # :60
        return loss.item(), iou.item(), pr5.item()                  # :68-70

    for _ in range(warmup):
        step()
    if on_gpu:
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        losses.append(step()[0])
    if on_gpu:
        e1.record()
        torch.cuda.synchronize(dev)
        secs = e0.elapsed_time(e1) / 1e3
    else:
        secs = time.perf_counter() - t0
    del model, opt
    return {"seconds": secs, "batch": batch, "losses": losses}


def eval_latency(arch: str, batch: int, iters: int, device: str = "cuda", amp_dtype=None):
    """tools/latency.py:51-66 on the reference model: eval, no_grad, sync per iteration; -> p50 ms."""
    from oracle import synth
    cfg, model, _ = build_reference_model(arch, dropout=0.0)
    dev = torch.device(device)
    if dev.type == "cuda":
        torch.backends.cudnn.benchmark = True
    model = model.to(dev).eval()
    size = 128 if arch == "tiny" else 416
    image = torch.randn(batch, 3, size, size, device=dev)
    text = torch.randint(1, synth.ARCHS[arch]["vocab"] - 2, size=(batch, cfg.word_len), device=dev).long()
    ts = []
    with torch.no_grad():
        for i in range(iters):
            t0 = time.perf_counter()
            with torch.autocast(dev.type, dtype=amp_dtype, enabled=amp_dtype is not None and dev.type == "cuda"):
                model(image, text)
            if dev.type == "cuda":
                torch.cuda.synchronize()
            if i >= iters // 5:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2]
