"""GPU: cris.pytorch_b200.optim.Adam (one kernel launch per parameter group) against torch.optim.Adam, the optimizer
the reference builds at train.py:105-107 and steps through GradScaler (engine/engine.py:52-57).
Tolerance: 5e-6 relative on parameters and moments after several steps (same fp32 arithmetic, fused differently)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
SHAPES = [(1,), (7,), (3, 5, 7), (8193,), (64, 3, 3, 3), (1000, 130), (5, 8192), (300001,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(*s, generator=g)).cuda()) for s in SHAPES]


def _groups(ps):
    return [{"params": ps[:4], "initial_lr": 1e-3}, {"params": ps[4:], "lr": 3e-3, "weight_decay": 0.01}]


def _close(a, b, tol=5e-6):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)) <= tol


def _set_grads(pa, pb, it, scale=1.0, poison=False):
    g = torch.Generator().manual_seed(100 + it)
    for a, b in zip(pa, pb):
        gr = torch.randn(*a.shape, generator=g).cuda() * scale
        if poison and a.numel() == 8193:
            gr.view(-1)[17] = float("inf")
        a.grad, b.grad = gr.clone(), gr.clone()


def test_adam_matches_torch_two_groups():
    from cris.pytorch_b200.optim import Adam
    pa, pb = _params(0), _params(0)
    oa = Adam(_groups(pa), lr=1e-3, weight_decay=0.0)
    ob = torch.optim.Adam(_groups(pb), lr=1e-3, weight_decay=0.0)
    sa = torch.optim.lr_scheduler.MultiStepLR(oa, milestones=[3], gamma=0.1)
    sb = torch.optim.lr_scheduler.MultiStepLR(ob, milestones=[3], gamma=0.1)
    for it in range(6):
        _set_grads(pa, pb, it)
        oa.step(); ob.step()
        sa.step(); sb.step()
        oa.zero_grad(); ob.zero_grad()
    for a, b in zip(pa, pb):
        assert _close(a.data, b.data)
        assert _close(oa.state[a]["exp_avg"], ob.state[b]["exp_avg"])
        assert _close(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"])
        assert float(oa.state[a]["step"]) == float(ob.state[b]["step"]) == 6.0


def test_adam_under_gradscaler_skips_inf_steps_and_interoperates_with_torch_state():
    from cris.pytorch_b200.optim import Adam
    pa, pb = _params(1), _params(1)
    oa, ob = Adam(_groups(pa), lr=2e-3), torch.optim.Adam(_groups(pb), lr=2e-3)
    sca, scb = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for sc in (sca, scb):
        sc.scale(torch.zeros(1, device="cuda"))  # lazily creates the scale tensor
    for it in range(5):
        assert sca.get_scale() == scb.get_scale()
        _set_grads(pa, pb, it, scale=sca.get_scale(), poison=(it == 2))  # gradients arrive scaled by the loss scale
        sca.step(oa); scb.step(ob)
        sca.update(); scb.update()
    assert sca.get_scale() == 512.0
    for a, b in zip(pa, pb):
        assert _close(a.data, b.data)
        assert float(oa.state[a]["step"]) == float(ob.state[b]["step"]) == 4.0  # the poisoned step did not count
    # checkpoints move between the two optimizers (train.py:166-170 resumes 'optimizer' state)
    pc = _params(1)
    for c, b in zip(pc, pb):
        c.data.copy_(b.data)
    oc = Adam(_groups(pc), lr=2e-3)
    import copy
    oc.load_state_dict(copy.deepcopy(ob.state_dict()))  # (load_state_dict keeps references to same-device tensors)
    _set_grads(pc, pb, 9)
    oc.step(); ob.step()
    for c, b in zip(pc, pb):
        assert _close(c.data, b.data)
    ob.load_state_dict(oc.state_dict())


def test_adam_refuses_cpu_parameters():
    from cris.pytorch_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        Adam([p]).step()
