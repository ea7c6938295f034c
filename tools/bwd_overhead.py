"""How much of the backward phase is NOT our captured graph?  Times (CUDA events, B=64 r50) the bare replay of the backward
graph against `scaler.scale(loss).backward()` (autograd + AccumulateGrad + anything it launches), and the host time of
the optimizer step.   python tools/bwd_overhead.py"""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg, model, groups = bench.build_model("r50", dropout=0.1)
    model = model.to(dev).train()
    from cris.pytorch_b200.optim import Adam
    opt = Adam(groups, lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    img, word, mask = synth.make_inputs(64, 0, 416, cfg.word_len, synth.ARCHS["r50"]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    res = []
    for it in range(8):
        pred, tgt, loss = model(img, word, mask)
        opt.zero_grad()
        torch.cuda.synchronize()
        ev[0].record()
        t0 = time.perf_counter()
        scaler.scale(loss).backward()
        t1 = time.perf_counter()
        ev[1].record()
        scaler.step(opt)
        scaler.update()
        t2 = time.perf_counter()
        ev[2].record()
        torch.cuda.synchronize()
        gs = next(iter(model._get_engine().graphs.values()))
        ev[3].record()
        gs.gb.replay()
        e4 = torch.cuda.Event(enable_timing=True)
        e4.record()
        torch.cuda.synchronize()
        if it >= 3:
            res.append((ev[0].elapsed_time(ev[1]), ev[3].elapsed_time(e4), ev[1].elapsed_time(ev[2]), 1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    n = len(res)
    a = [sum(r[i] for r in res) / n for i in range(5)]
    print(f"loss.backward() GPU {a[0]:.2f} ms | bare backward-graph replay {a[1]:.2f} ms | difference {a[0] - a[1]:.2f} ms")
    print(f"optimizer phase GPU {a[2]:.2f} ms | host: backward() call {a[3]:.2f} ms, scaler.step+update {a[4]:.2f} ms")


if __name__ == "__main__":
    main()
