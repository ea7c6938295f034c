"""GPU: parity of the sm_100a path on the configurations the metric is quoted on, against the committed outputs of
the UNMODIFIED reference (tests/golden/*.pt, oracle/make_golden.py):

  * cris_r50 B=2 and B=8, cris_r101 B=4 at 416x416 / 17 tokens: eval logits, thresholded mask, train loss, the
    nearest-resized mask, and EVERY parameter's gradient (norm + the 24 stored samples), grouped per block;
  * per-stage intermediates (stem, layer1-4, attnpool, word/state, fq, dec0-2, proj_feat) against the fp32 oracle;
  * 2 ranks x B=4 with SyncBatchNorm + DistributedDataParallel == 1 GPU x B=8 == the B=8 golden
    (reference train.py:97-102: SyncBN makes the statistics global, DDP averages the gradients).

Stated tolerances (bf16 storage / tensor-core operands with fp32 accumulation vs the fp32 reference):
  eval logits  rel L2 <= 3e-2; a mask pixel may differ only if its reference logit is within 0.05 of the threshold
  train loss   |delta| <= 3e-2 (batch-statistics BatchNorm amplifies storage noise; see oracle/synth.py)
  gradients    per block (norm of the block's gradient, cosine of its concatenated samples vs the reference):
                 image tower (stem, layer1-4, attnpool): norm ratio in [0.9, 1.1], cosine >= 0.55
                 text tower, neck, decoder, projector  : norm ratio in [0.9, 1.1], cosine >= 0.88
               (floors sit below the spread of five full runs on different boxes: image-tower blocks 0.67-0.91, worst
               r101 layer2 0.675 / 0.706 / 0.727 / 0.746 / 0.807; r101 text blocks 0-5 0.940-0.971; neck 0.93-0.98)
               per parameter: norm ratio in [0.8, 1.25] for >= 97 % of the tensors
               Measured (round 2, gpurun_out/parity_*.json): r50 B=8 every tensor within [0.87, 1.12], head / text /
               neck cosines 0.98-1.00, image-tower cosines 0.80-0.90 (r101, 101 layers deep: 0.68-0.79).  That is the
               RUN-TO-RUN floor of a bf16-storage pass, not an arithmetic error: two executions of the same
               single-GPU configuration that differ only in the arrival order of fp32 atomics disagree by the same
               amount (measured in tests/test_syncbn_equiv_gpu.py) — a 1e-7 perturbation flips a few bf16 roundings and
               within ~4 layers the two trajectories differ by the bf16 rounding noise, which flips ~0.4 % of the
               ReLU masks per layer.  The per-kernel backward checks at 1-4 % live in tests/test_ops_gpu.py.
  B=2 caveat   neck.txt_proj is BatchNorm1d over the BATCH: with 2 samples its output is +-gamma+beta whatever the
               input, the true gradient through it is ~0 (an eps effect) and is amplified by invstd up to 316x, so
               at B=2 everything upstream of it (text tower, neck.txt_proj.0) is ill-conditioned in ANY precision.
               Those blocks are compared at B=8 (r50) and B=4 (r101) only.
A JSON report with every number is written to gpurun_out/parity_<tag>.json on each run.
"""
import json
import os

import pytest
import torch

from oracle import cris_oracle as O
from oracle import synth
from oracle.hostinfo import usable_cpus

from parity_util import REPO, THR, build, gradient_report, group_of, rel

pytestmark = pytest.mark.gpu
PARAM_OUTLIERS = 0.03
ILL_CONDITIONED_AT_B2 = ("backbone.text.", "neck.txt_proj")  # upstream of BatchNorm1d over a batch of 2


def _is_image_tower(group: str) -> bool:
    return group.startswith("backbone.visual")


def check_against_golden(arch, tag, golden_dir, B):
    g = torch.load(os.path.join(golden_dir, tag + ".pt"), weights_only=False)
    cfg, sd, model = build(arch)
    img, word, mask = synth.make_inputs(B, 0, 416, cfg.word_len, synth.ARCHS[arch]["vocab"])
    rep = {"tag": tag}
    model.eval()
    with torch.no_grad():
        pred = model(img.cuda(), word.cuda()).cpu()
    ref = g["eval_pred"]
    assert pred.shape == ref.shape == (B, 1, 104, 104)
    rep["eval_rel"] = rel(pred, ref)
    flips = (pred > THR) != (ref > THR)
    band = (ref - THR).abs() <= 0.05
    rep["flips"] = int(flips.sum())
    rep["flips_outside_band"] = int((flips & ~band).sum())
    rep["in_band_fraction"] = float(band.float().mean())
    a, b = pred > THR, ref > THR
    rep["iou_vs_reference_mask"] = float((a & b).sum()) / max(1.0, float((a | b).sum()))
    model.train()
    model.zero_grad(set_to_none=True)
    p2, m2, loss = model(img.cuda(), word.cuda(), mask.cuda())
    rep["mask_equal"] = bool(torch.equal(m2.cpu(), g["train_mask"]))
    rep["loss"], rep["loss_ref"] = float(loss), float(g["train_loss"])
    rep["train_pred_rel"] = rel(p2.cpu(), g["train_pred"])
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == {k for k, v in g["grads"].items() if v is not None}
    skip = ILL_CONDITIONED_AT_B2 if B < 4 else ()
    per, groups = gradient_report(grads, g["grads"], skip_prefixes=skip)
    rep["groups"] = groups
    ratios = sorted(((abs(v["ratio"] - 1.0), k, v["ratio"], v["sample_rel"]) for k, v in per.items()
                     if v["ratio"] is not None and v["norm_ref"] >= 1e-7 and "txt_proj.1" not in k
                     and not any(group_of(k).startswith(p) or k.startswith(p) for p in skip)), reverse=True)
    rep["worst10"] = [{"name": k, "ratio": r, "sample_rel": s} for _, k, r, s in ratios[:10]]
    rep["param_outlier_fraction"] = sum(1 for _, _, r, _ in ratios if not (0.8 <= r <= 1.25)) / max(1, len(ratios))
    msd = model.state_dict()
    run_err = []
    for k, v in g["running"].items():
        if "txt_proj" in k:
            continue
        n = float(msd[k].double().norm())
        run_err.append(abs(n - float(v["norm"])) / max(float(v["norm"]), 1e-6))
    rep["running_stat_norm_err_max"] = max(run_err)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(REPO, "gpurun_out", f"parity_{tag}.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in rep.items() if k not in ("groups",)}))
    for name, G in sorted(groups.items()):
        print(f"  {name:34s} norm_ratio {G['norm_ratio']:.3f} cos {G['cos']:.3f} ({G['tensors']} tensors)")
    # ---- assertions (tolerances stated in the module docstring) ----
    assert rep["eval_rel"] <= 3e-2
    assert rep["flips_outside_band"] == 0
    assert rep["iou_vs_reference_mask"] >= 0.93
    assert rep["mask_equal"]
    assert abs(rep["loss"] - rep["loss_ref"]) <= 3e-2
    assert rep["running_stat_norm_err_max"] <= 5e-2
    for name, G in groups.items():
        assert 0.9 <= G["norm_ratio"] <= 1.1, (name, G)
        floor = 0.55 if _is_image_tower(name) else 0.88
        assert G["cos"] >= floor, (name, G, floor)
    assert rep["param_outlier_fraction"] <= PARAM_OUTLIERS, rep["worst10"]
    return rep


def test_r50_b2_every_parameter_gradient(golden_dir):
    check_against_golden("r50", "r50_b2_416", golden_dir, 2)


def test_r50_b8_matches_reference_golden(golden_dir):
    check_against_golden("r50", "r50_b8_416", golden_dir, 8)


def test_r101_matches_reference_golden(golden_dir):
    check_against_golden("r101", "r101_b4_416", golden_dir, 4)


# per-stage tolerances: eval = bf16 storage vs fp32 (error grows slowly with depth); train = vs the oracle with the
# engine's bf16 storage points emulated (batch statistics amplify the remaining ordering noise)
EVAL_TAP_TOL = {"stem": 8e-3, "layer1": 1.2e-2, "layer2": 1.2e-2, "layer3": 1.2e-2, "layer4": 1.2e-2, "attnpool": 1.5e-2,
                "word": 1.5e-2, "state": 1.5e-2, "fq": 2.5e-2, "dec0": 2.5e-2, "dec1": 2.5e-2, "dec2": 2.5e-2,
                "dec_out": 2.5e-2, "proj_feat": 2.5e-2}
TRAIN_TAP_TOL = {"stem": 2e-3, "layer1": 5e-3, "layer2": 1.5e-2, "layer3": 3e-2, "layer4": 4e-2, "attnpool": 5e-2,
                 "word": 2e-3, "state": 2e-3, "fq": 0.12, "dec0": 0.12, "dec1": 0.12, "dec2": 0.12, "dec_out": 0.12,
                 "proj_feat": 0.15}


def test_per_stage_intermediates_tiny():
    """Every tapped stage of the forward against the oracle (tools/parity_report.py as a test)."""
    torch.set_num_threads(min(16, usable_cpus()))
    cfg, sd, model = build("tiny")
    eng = model._get_engine()
    img, word, mask = synth.make_inputs(2, 0, 128, cfg.word_len, synth.ARCHS["tiny"]["vocab"])
    otaps = {}
    with torch.no_grad():
        O.cris_forward(sd, img, word, training=False, num_head=cfg.num_head, taps=otaps)
    model.eval()
    eng.debug_taps = {}
    try:
        with torch.no_grad():
            model(img.cuda(), word.cuda())
        got = dict(eng.debug_taps)
        eng.debug_taps = {}
        ttaps = {}
        with torch.no_grad():
            O.cris_forward(sd, img, word, mask, training=True, num_head=cfg.num_head, taps=ttaps, storage="bf16")
        model.train()
        with torch.no_grad():
            model(img.cuda(), word.cuda(), mask.cuda())
        got_train = dict(eng.debug_taps)
    finally:
        eng.debug_taps = None
    assert set(EVAL_TAP_TOL) <= set(got) and set(EVAL_TAP_TOL) <= set(otaps)
    errs = {}
    for k, tol in EVAL_TAP_TOL.items():
        o = otaps[k]
        v = got[k].cpu()
        if v.dim() == 2 and o.dim() == 3:
            o = o.reshape(-1, o.shape[-1])
        errs["eval." + k] = (rel(v, o), tol)
    for k, tol in TRAIN_TAP_TOL.items():
        o = ttaps[k].detach()
        v = got_train[k].cpu()
        if v.dim() == 2 and o.dim() == 3:
            o = o.reshape(-1, o.shape[-1])
        errs["train." + k] = (rel(v, o), tol)
    print({k: round(e, 5) for k, (e, _) in errs.items()})
    bad = {k: (e, t) for k, (e, t) in errs.items() if e > t}
    assert not bad, bad
