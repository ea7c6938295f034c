"""Run a few cris_r50 train steps for profiling: warm-up steps outside, ONE step inside cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [batch]
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    os.environ.setdefault("CRIS_B200_GRAPHS", "0")  # per-kernel launch list: eager launches
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    arch = sys.argv[2] if len(sys.argv) > 2 else "r50"
    dev = torch.device("cuda", 0)
    cfg, model, groups = bench.build_model(arch, dropout=0.1)
    model = model.to(dev).train()
    from cris.pytorch_b200.optim import Adam
    opt = Adam(groups, lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    size = 416 if arch != "tiny" else 128
    img, word, mask = synth.make_inputs(B, 0, size, cfg.word_len, synth.ARCHS[arch]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)

    def step():
        pred, tgt, loss = model(img, word, mask)
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        bench.train_metric(pred, tgt)
        return loss.item()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    eng = model._get_engine()
    eng.gemm_log = []
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    import json
    json.dump(eng.gemm_log, open(os.path.join(REPO, 'gpurun_out', 'gemm_log.json'), 'w'))


if __name__ == "__main__":
    main()
