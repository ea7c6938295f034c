"""Per-phase timing of one cris_r50 train step (CUDA events): forward, backward, optimizer, metrics."""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda", 0)
    cfg, model, groups = bench.build_model("r50", dropout=0.1)
    model = model.to(dev).train()
    from cris.pytorch_b200.optim import Adam
    opt = (torch.optim.Adam if os.environ.get("CRIS_B200_TORCH_ADAM") == "1" else Adam)(groups, lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    img, word, mask = synth.make_inputs(B, 0, 416, cfg.word_len, synth.ARCHS["r50"]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for it in range(8):
        t0 = time.perf_counter()
        ev[0].record()
        pred, tgt, loss = model(img, word, mask)
        ev[1].record()
        t1 = time.perf_counter()
        opt.zero_grad()
        scaler.scale(loss).backward()
        ev[2].record()
        t2 = time.perf_counter()
        scaler.step(opt)
        scaler.update()
        ev[3].record()
        t3 = time.perf_counter()
        bench.train_metric(pred, tgt)
        l = loss.item()
        ev[4].record()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if it >= 3:
            print(f"it {it}: gpu fwd {ev[0].elapsed_time(ev[1]):7.2f} bwd {ev[1].elapsed_time(ev[2]):7.2f} opt {ev[2].elapsed_time(ev[3]):7.2f} "
                  f"metric {ev[3].elapsed_time(ev[4]):6.2f} | host fwd {1e3*(t1-t0):6.2f} bwd {1e3*(t2-t1):6.2f} opt {1e3*(t3-t2):6.2f} rest {1e3*(t4-t3):6.2f} | total {1e3*(t4-t0):7.2f} ms loss {l:.4f}")
    print("adam table refreshes", getattr(opt, "table_refreshes", None))
    print("mem GB", torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)


if __name__ == "__main__":
    main()
