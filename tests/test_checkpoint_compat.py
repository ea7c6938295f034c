"""CPU: on-disk checkpoint compatibility with the reference (train.py:160-176,192-207, test.py:71-79).

With the unmodified reference staged under baseline/_ref (tools/stage_reference.py; skipped where it is absent):
  * a `last_model.pth` written the reference's way from the REFERENCE model (DataParallel-prefixed state_dict, Adam,
    MultiStepLR) loads strictly into cris.pytorch_b200.CRIS, and
  * the same file written from OUR module loads strictly into the reference's `model.segmenter.CRIS` through the
    `DataParallel(...).load_state_dict(strict=True)` line of test.py — values bit-identical both ways;
  * torch.optim.Adam state written by the reference resumes in cris.pytorch_b200.optim.Adam's state layout and back
    (the optimizer classes share param-group order and per-parameter `step / exp_avg / exp_avg_sq` keys).
"""
import os
import sys
import tempfile

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))


def _both(arch="tiny"):
    from baseline import ref_step
    if not ref_step.available():
        pytest.skip("baseline/_ref is not staged")
    from cris.pytorch_b200 import build_segmenter
    from oracle import synth
    cfg_r, ref_model, ref_groups = ref_step.build_reference_model(arch)
    cfg = synth.make_cfg(arch)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 1), p)   # different seed: loading must overwrite it
        cfg.clip_pretrain = p
        ours, our_groups = build_segmenter(cfg)
    return ref_model, ref_groups, ours, our_groups


def test_checkpoint_round_trip_with_the_reference_model():
    import checkpoint_tool as T
    from torch.optim.lr_scheduler import MultiStepLR
    ref_model, ref_groups, ours, our_groups = _both()
    opt = torch.optim.Adam(ref_groups, lr=1e-4)
    sch = MultiStepLR(opt, milestones=[35], gamma=0.1)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "last_model.pth")
        T.save_checkpoint(path, torch.nn.DataParallel(ref_model), opt, sch, 3, 0.5, 0.6, {"Pr@50": 0.7})   # train.py:192-203
        ck = torch.load(path, map_location="cpu", weights_only=False)
        assert list(ck.keys()) == list(T.CKPT_KEYS) and all(k.startswith("module.") for k in ck["state_dict"])
        torch.nn.DataParallel(ours).load_state_dict(ck["state_dict"], strict=True)                         # test.py:71-78
        rs = ref_model.state_dict()
        for k, v in ours.state_dict().items():
            assert torch.equal(v, rs[k]), k
        # ... and back: our module's checkpoint into the reference class
        with torch.no_grad():
            for p in ours.parameters():
                p.add_(0.01)
        path2 = os.path.join(td, "best_model.pth")
        our_opt = torch.optim.Adam(our_groups, lr=1e-4)
        T.save_checkpoint(path2, torch.nn.DataParallel(ours), our_opt, MultiStepLR(our_opt, [35], 0.1), 4, 0.5, 0.6, {})
        ck2 = torch.load(path2, map_location="cpu", weights_only=False)
        torch.nn.DataParallel(ref_model).load_state_dict(ck2["state_dict"], strict=True)
        os_ = ours.state_dict()
        assert list(ref_model.state_dict().keys()) == list(os_.keys())
        for k, v in ref_model.state_dict().items():
            assert torch.equal(v, os_[k]), k
        assert T.inspect(ck2)["module_prefix"] and T.inspect(ck2)["optimizer_groups"] == [len(g["params"]) for g in our_groups]
        assert T.verify(ck2, "tiny")


def test_optimizer_state_layout_is_interchangeable():
    """torch.optim.Adam.state_dict() <-> cris.pytorch_b200.optim.Adam: same groups, same per-parameter keys (the CUDA
    step itself is covered by tests/test_optim_gpu.py; this is the resume path of train.py:166-170 on the CPU)."""
    from cris.pytorch_b200.optim import Adam
    ref_model, ref_groups, ours, our_groups = _both()
    t = torch.optim.Adam(ref_groups, lr=1e-4)
    for g in t.param_groups:
        for p in g["params"]:
            p.grad = torch.zeros_like(p) + 1e-3
    t.step()
    sd = t.state_dict()
    mine = Adam(our_groups, lr=1e-4)
    mine.load_state_dict(sd)
    back = mine.state_dict()
    assert [len(g["params"]) for g in back["param_groups"]] == [len(g["params"]) for g in sd["param_groups"]]
    assert back["state"].keys() == sd["state"].keys()
    k0 = next(iter(sd["state"]))
    assert set(back["state"][k0].keys()) == set(sd["state"][k0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    torch.optim.Adam(ref_groups, lr=1e-4).load_state_dict(back)   # and the reference optimizer takes it back
