"""GPU: N-GPU == 1-GPU-global-batch.  Two ranks x B=4 with SyncBatchNorm + DistributedDataParallel (the reference's
train.py:97-102 recipe) must give the loss and parameter gradients of ONE process running the same 8 samples —
SyncBN makes the batch statistics global, DDP averages the per-rank mean losses' gradients — and both must agree
with the committed B=8 outputs of the unmodified reference (tests/golden/r50_b8_416.pt).

On a box with >= 2 GPUs the ranks sit on different devices; on a single-GPU box both use cuda:0 (the driver
time-slices the two contexts; the NVLink peer-exchange kernels then hand over through the same IPC mapping), like
tests/test_peer_gpu.py.  The process group is gloo so that the test does not depend on two NCCL devices.

What "equal" can mean here.  Activations are STORED in bf16: a perturbation of 1e-7 (the fp32 summation order of a
BatchNorm statistic: rank-ordered partial sums, or merely the arrival order of fp32 atomics) flips a few bf16
roundings, each flip is a 0.4 % change of one element, and within about four layers ANY two non-bit-identical
executions differ by the bf16 rounding noise itself; downstream ReLU masks then differ in ~0.4 % of the units per
layer.  So two runs of the SAME single-GPU configuration already disagree in the image tower's gradient direction
(cosine 0.8-0.9, measured below as the run-to-run floor) while norms, the loss and every other block agree closely.
The test therefore measures that floor (1x8 run twice) and requires the 2x4 SyncBN/DDP run to sit at it:
  loss |delta| <= 2e-3 (and <= 3e-2 vs the fp32 reference); every gradient tensor's norm within 5 % for >= 97 % of
  the tensors; per block, cosine(2x4, 1x8) >= cosine(1x8, 1x8') - 0.08 and >= 0.62 (image tower) / 0.92 (rest).
"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth
from parity_util import build, group_of

pytestmark = pytest.mark.gpu
ARCH, B_GLOBAL, SIZE = "r50", 8, 416


def _summary(named_grads):
    out = {}
    for k, g in named_grads.items():
        f = g.detach().flatten()
        gen = torch.Generator().manual_seed(f.numel() % 100003)
        idx = torch.randint(0, f.numel(), (min(24, f.numel()),), generator=gen)
        # plain Python numbers: the summary crosses a multiprocessing queue (no shared tensor storage to keep alive)
        out[k] = {"norm": float(f.double().norm()), "numel": f.numel(), "val": f[idx.to(f.device)].double().cpu().tolist()}
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CRIS_B200_PEER_TIMEOUT_S="120")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        cfg, sd, model = build(ARCH)
        model = model.to(dev)
        eng = model._get_engine()
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)          # train.py:97-98
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)
        img, word, mask = synth.make_inputs(B_GLOBAL, 0, SIZE, cfg.word_len, synth.ARCHS[ARCH]["vocab"])
        per = B_GLOBAL // world
        sl = slice(rank * per, (rank + 1) * per)
        ddp.train()
        pred, m, loss = ddp(img[sl].to(dev), word[sl].to(dev), mask[sl].to(dev))
        loss.backward()
        torch.cuda.synchronize(dev)
        ld = loss.detach().clone().cpu()
        dist.all_reduce(ld)
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        used_peer = eng.peer_exchange(dev) is not None
        # running statistics must be bit-identical on every rank (DDP's per-forward buffer broadcast is skipped for them:
        # CRIS._ddp_params_and_buffers_to_ignore)
        stats = torch.cat([b.detach().double().reshape(-1).cpu() for k, b in model.named_buffers() if "running_" in k])
        digest = (float(stats.sum()), float((stats * stats).sum()), float(stats.abs().max()))
        q.put((rank, "ok", float(ld) / world, _summary(grads) if rank == 0 else None, used_peer, bool(eng.graphs), digest))
        dist.barrier()
    except Exception as e:  # surface worker failures instead of a queue timeout
        q.put((rank, "error", repr(e), None, False, False, None))
        raise
    finally:
        dist.destroy_process_group()


def _group_cos(sa, sb):
    """per-block cosine / norm ratio of two gradient summaries (samples weighted by tensor size)"""
    G = {}
    for k, a in sa.items():
        if a["norm"] < 1e-7 or "txt_proj.1" in k:
            continue
        b = sb[k]
        va, vb = torch.tensor(a["val"]), torch.tensor(b["val"])
        w = a["numel"] / max(1, va.numel())
        g = G.setdefault(group_of(k), [0.0, 0.0, 0.0, 0.0, 0.0])
        g[0] += w * float(va @ vb); g[1] += w * float(va @ va); g[2] += w * float(vb @ vb)
        g[3] += a["norm"] ** 2; g[4] += b["norm"] ** 2
    return {k: {"cos": g[0] / ((g[1] * g[2]) ** 0.5 + 1e-30), "norm_ratio": (g[4] / g[3]) ** 0.5} for k, g in G.items()}


def _single_run(seed_note=""):
    cfg, sd, model = build(ARCH)
    img, word, mask = synth.make_inputs(B_GLOBAL, 0, SIZE, cfg.word_len, synth.ARCHS[ARCH]["vocab"])
    model.train()
    _, _, loss = model(img.cuda(), word.cuda(), mask.cuda())
    loss.backward()
    out = _summary({k: p.grad for k, p in model.named_parameters() if p.grad is not None}), float(loss)
    del model
    torch.cuda.empty_cache()
    return out


def test_two_ranks_syncbn_ddp_equal_one_gpu_global_batch(golden_dir):
    # 1) one process, the whole batch — twice: the second run measures the run-to-run floor of a bf16-storage pass
    single, loss1 = _single_run()
    again, loss1b = _single_run()
    floor = _group_cos(single, again)
    # 2) two ranks, half the batch each
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(r[1] == "ok" for r in res), res
    assert res[0][6] == res[1][6], ("BatchNorm running statistics differ between the ranks", res[0][6], res[1][6])
    r0 = next(r for r in res if r[0] == 0)
    loss2, multi = r0[2], r0[3]
    print(f"loss 1x8 {loss1:.6f} / {loss1b:.6f}  2x4 {loss2:.6f}  peer exchange {r0[4]}  graphs {r0[5]}")
    assert abs(loss1 - loss2) <= 2e-3 and abs(loss1 - loss1b) <= 2e-3
    bad = [(k, multi[k]["norm"] / s["norm"]) for k, s in single.items()
           if s["norm"] >= 1e-7 and abs(multi[k]["norm"] / s["norm"] - 1.0) > 0.05]
    # the same count between the two single-GPU runs = the run-to-run floor of this statistic (stem / layer1 BatchNorm
    # parameters sit at the end of the longest bf16 chain; observed 9-14 of 448 for 2x4 vs 1x8 across boxes)
    floor_bad = [k for k, s in single.items()
                 if s["norm"] >= 1e-7 and abs(again[k]["norm"] / s["norm"] - 1.0) > 0.05]
    print(f"{len(bad)} of {len(single)} gradient norms differ by more than 5 % (between the two 1x8 runs: {len(floor_bad)}): {bad[:5]}")
    assert len(bad) <= max(0.05 * len(single), 3 * len(floor_bad)), bad[:10]
    assert all(abs(r - 1.0) < 0.35 for _, r in bad), bad[:10]
    got = _group_cos(single, multi)
    for name in sorted(got):
        print(f"  {name:34s} cos(2x4, 1x8) {got[name]['cos']:.3f}   run-to-run floor cos(1x8, 1x8') {floor[name]['cos']:.3f}"
              f"   norm ratio {got[name]['norm_ratio']:.3f}")
    for name, g in got.items():
        lo = 0.62 if name.startswith("backbone.visual") else 0.92
        assert g["cos"] >= lo and g["cos"] >= floor[name]["cos"] - 0.08, (name, g, floor[name])
        assert 0.95 <= g["norm_ratio"] <= 1.05, (name, g)
    # 3) both agree with the unmodified reference at B=8
    g = torch.load(os.path.join(golden_dir, "r50_b8_416.pt"), weights_only=False)
    assert abs(loss2 - float(g["train_loss"])) <= 3e-2
    ok = total = 0
    for k, gg in g["grads"].items():
        if gg is None or float(gg["norm"]) < 1e-7 or "txt_proj.1" in k:
            continue
        total += 1
        ok += int(0.8 <= multi[k]["norm"] / float(gg["norm"]) <= 1.25)
    assert ok >= 0.97 * total, (ok, total)
