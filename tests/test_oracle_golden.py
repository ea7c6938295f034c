"""CPU: the oracle restatement (oracle/cris_oracle.py) replays the committed outputs of the
UNMODIFIED reference (tests/golden/*.pt, made by oracle/make_golden.py from /root/reference)."""
import os

import pytest
import torch

from oracle import cris_oracle as O
from oracle import synth


def _load(golden_dir, tag):
    return torch.load(os.path.join(golden_dir, tag + ".pt"), weights_only=False)


@pytest.mark.parametrize("tag,arch,check_grads", [("tiny_b2_128", "tiny", True), ("r50_b2_416", "r50", False),
                                                  ("r101_b4_416", "r101", False), ("r50_b8_416", "r50", False)])
def test_oracle_matches_reference_outputs(golden_dir, tag, arch, check_grads):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g = _load(golden_dir, tag)
    cfg = synth.make_cfg(arch)
    sd = synth.full_state_dict(arch, 0, cfg)
    img, word, mask = synth.make_inputs(g["batch"], 0, g["size"], cfg.word_len, synth.ARCHS[arch]["vocab"])
    with torch.no_grad():
        ev = O.cris_forward(sd, img, word, training=False, num_head=cfg.num_head)["pred"]
    scale = g["eval_pred"].abs().max().item()
    assert (ev - g["eval_pred"]).abs().max().item() <= 2e-4 * scale  # fp32 CPU vs fp32 CPU
    # thresholded mask (sigmoid > 0.35  <=>  logit > -0.619) must agree exactly except for logits that sit
    # within fp32 rounding distance of the threshold
    thr = -0.6190392
    flips = ((ev > thr) != (g["eval_pred"] > thr)) & ((g["eval_pred"] - thr).abs() > 1e-3)
    assert int(flips.sum()) == 0

    if check_grads:
        sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
               for k, v in sd.items()}
    else:
        sdg = sd
    ctx = torch.enable_grad() if check_grads else torch.no_grad()
    with ctx:
        tr = O.cris_forward(sdg, img, word, mask, training=True, num_head=cfg.num_head)
    assert (tr["pred"].detach() - g["train_pred"]).abs().max().item() <= 5e-4 * scale
    assert torch.equal(tr["mask"], g["train_mask"])
    assert abs(tr["loss"].item() - g["train_loss"].item()) <= 1e-5
    for k, v in g["running"].items():
        nv = tr["new_running"][k]
        assert abs(nv.double().norm().item() - v["norm"].item()) <= 1e-4 * max(v["norm"].item(), 1e-6), k
    if check_grads:
        tr["loss"].backward()
        for k, gg in g["grads"].items():
            if gg is None:  # backbone.logit_scale never receives a gradient
                assert sdg[k].grad is None, k
                continue
            gr = sdg[k].grad.flatten()
            if gg["norm"].item() < 1e-7:  # analytically-zero gradients (e.g. attnpool.k_proj.bias)
                assert gr.double().norm().item() < 1e-6, k
                continue
            assert abs(gr.double().norm().item() - gg["norm"].item()) <= 1e-2 * gg["norm"].item(), k
            assert (gr[gg["idx"]] - gg["val"]).abs().max().item() <= 3e-2 * gg["val"].abs().max().item() + 1e-9, k


def test_reference_constructor_contract(golden_dir):
    """What the reference ctor does to the checkpoint (model/clip.py:477-500,552-554 +
    segmenter.py:16): conv/linear/MHA/text_projection tensors are fp16-rounded, the rest untouched."""
    g = _load(golden_dir, "tiny_b2_128")["ctor"]
    rounded_expected = lambda k: (k.endswith("conv1.weight") or k.endswith("conv2.weight") or k.endswith("conv3.weight")  # noqa: E731
                                  or "downsample.0" in k or "_proj." in k or "in_proj_" in k or ".mlp.c_" in k
                                  or k.endswith("text_projection"))
    for k, unchanged in g["unchanged"].items():
        if rounded_expected(k):
            assert g["fp16_rounded"][k], k
        else:
            assert unchanged, k
    assert all(d == "torch.float32" or k.endswith("num_batches_tracked") for k, d in g["dtypes"].items())
