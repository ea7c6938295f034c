"""Diagnostic: does the 177 MB input copy overlap the train step? Times (CUDA events) the copy alone, the step alone,
and both issued together (copy on a side stream)."""
import os
import sys

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
    img_h, word_h, mask_h = img.pin_memory(), word.pin_memory(), mask.pin_memory()
    img_d, word_d, mask_d = img_h.to(dev), word_h.to(dev), mask_h.to(dev)
    side = torch.cuda.Stream()

    def step():
        pred, tgt, loss = model(img_d, word_d, mask_d)
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()

    def copy(stream):
        with torch.cuda.stream(stream):
            a = img_h.to(dev, non_blocking=True); b = word_h.to(dev, non_blocking=True); c = mask_h.to(dev, non_blocking=True)
        return a, b, c

    def ev():
        return torch.cuda.Event(enable_timing=True)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    main_s = torch.cuda.current_stream()
    for rep in range(3):
        a0, a1 = ev(), ev()
        a0.record(main_s); keep = copy(main_s); a1.record(main_s); torch.cuda.synchronize()
        s0, s1 = ev(), ev()
        s0.record(); step(); s1.record(); torch.cuda.synchronize()
        c0, c1, m0, m1 = ev(), ev(), ev(), ev()
        m0.record(main_s)
        c0.record(side); keep2 = copy(side); c1.record(side)
        step()
        m1.record(main_s)
        torch.cuda.synchronize()
        print(f"rep {rep}: copy alone {a0.elapsed_time(a1):.2f} ms | step alone {s0.elapsed_time(s1):.2f} ms | "
              f"together: copy {c0.elapsed_time(c1):.2f} ms, step {m0.elapsed_time(m1):.2f} ms", flush=True)


if __name__ == "__main__":
    main()
