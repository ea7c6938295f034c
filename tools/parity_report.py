"""GPU parity report: cris.pytorch_b200 (sm_100a kernels) vs the CPU fp32 oracle, stage by stage.

    python tools/parity_report.py [tiny|r50] [batch] [size]

Test/debug tool (it imports oracle/), prints relative errors of every tapped intermediate, of the logits,
the loss and every parameter gradient.  Runs on the GPU box; writes nothing.
"""
import os
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import cris_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402
from cris.pytorch_b200 import CRIS  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(arch, dropout=0.0):
    cfg = synth.make_cfg(arch, dropout=dropout)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        cfg.clip_pretrain = path
        model = CRIS(cfg)
    sd = synth.full_state_dict(arch, 0, cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    size = int(sys.argv[3]) if len(sys.argv) > 3 else (128 if arch == "tiny" else 416)
    from oracle.hostinfo import usable_cpus
    torch.set_num_threads(min(32, usable_cpus()))
    print(f"[host] usable cpus {usable_cpus()} (os.cpu_count {os.cpu_count()})")
    cfg, sd, model = build(arch)
    img, word, mask = synth.make_inputs(B, 0, size, cfg.word_len, synth.ARCHS[arch]["vocab"])
    model = model.cuda()
    eng = model._get_engine()
    # ---------------- eval ----------------
    otaps = {}
    with torch.no_grad():
        ref = O.cris_forward(sd, img, word, training=False, num_head=cfg.num_head, taps=otaps)
    model.eval()
    eng.debug_taps = {}
    t0 = time.time()
    with torch.no_grad():
        pred = model(img.cuda(), word.cuda())
    torch.cuda.synchronize()
    print(f"[eval] forward ok in {time.time() - t0:.2f}s; launches so far {__import__('cris.pytorch_b200._lib', fromlist=['x']).launch_count()}")
    for k, v in eng.debug_taps.items():
        if k in otaps:
            o = otaps[k]
            vv = v.cpu()
            if vv.dim() == 2 and o.dim() == 3:
                o = o.reshape(-1, o.shape[-1])
            print(f"[eval] tap {k:10s} rel err {rel(vv, o):.4e}   (|ref| {o.abs().mean():.3f})")
    p = pred.cpu()
    thr = -0.6190392
    flips = int(((p > thr) != (ref['pred'] > thr)).sum())
    print(f"[eval] pred rel err {rel(p, ref['pred']):.4e} max|d| {(p - ref['pred']).abs().max():.4f} "
          f"std(ref) {ref['pred'].std():.4f} mask flips {flips}/{p.numel()}")
    # ---------------- train ----------------
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
           for k, v in sd.items()}
    otaps = {}
    t0 = time.time()
    storage = os.environ.get("PARITY_STORAGE", "bf16")
    ref = O.cris_forward(sdg, img, word, mask, training=True, num_head=cfg.num_head, taps=otaps, storage=storage)
    ref["loss"].backward()
    print(f"[train] oracle ({storage} storage) fwd+bwd {time.time() - t0:.1f}s")
    model.train()
    eng.debug_taps = {}
    pred, m, loss = model(img.cuda(), word.cuda(), mask.cuda())
    for k, v in eng.debug_taps.items():
        if k in otaps:
            o = otaps[k].detach()
            vv = v.cpu()
            if vv.dim() == 2 and o.dim() == 3:
                o = o.reshape(-1, o.shape[-1])
            print(f"[train] tap {k:10s} rel err {rel(vv, o):.4e}")
    eng.debug_taps = None
    print(f"[train] loss {float(loss):.6f} ref {float(ref['loss']):.6f}  pred rel err {rel(pred.cpu(), ref['pred'].detach()):.4e} "
          f"mask equal {bool(torch.equal(m.cpu(), ref['mask']))}")
    loss.backward()
    torch.cuda.synchronize()
    worst = []
    for k, prm in model.named_parameters():
        g_ref = sdg[k].grad
        if prm.grad is None:
            print(f"[grad] {k}: no grad (ref {'None' if g_ref is None else 'present'})")
            continue
        r = rel(prm.grad.cpu(), g_ref)
        worst.append((r, k, float(g_ref.norm())))
    worst.sort(reverse=True)
    for r, k, n in (worst if os.environ.get("PARITY_ALL") else worst[:25]):
        print(f"[grad] {r:.3e}  |ref| {n:.3e}  {k}")
    import statistics
    print(f"[grad] median rel err {statistics.median([w[0] for w in worst]):.3e} over {len(worst)} tensors")
    for k, v in ref["new_running"].items():
        r = rel(model.state_dict()[k].cpu(), v)
        if r > 2e-2:
            print(f"[running] {k} rel err {r:.3e}")
    print("[done]")


if __name__ == "__main__":
    main()
