"""Debug aid: gradients of the tiny model with the weight-gradient branch off / on, every parameter, two replays each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(arch, dropout=0.0):
    import tempfile
    from cris.pytorch_b200 import CRIS
    cfg = synth.make_cfg(arch, dropout=dropout)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        cfg.clip_pretrain = path
        model = CRIS(cfg)
    sd = synth.full_state_dict(arch, 0, cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.cuda()


def main():
    cfg, sd, model = build("tiny")
    eng = model._get_engine()
    img, word, mask = synth.make_inputs(2, 5, 128, cfg.word_len, synth.ARCHS["tiny"]["vocab"])
    img, word, mask = img.cuda(), word.cuda(), mask.cuda()
    model.train()
    saved = {k: b.clone() for k, b in model.named_buffers()}
    res = {}
    for tag, env in (("off", {"CRIS_B200_WGRAD_STREAM": "0"}), ("off2", {"CRIS_B200_WGRAD_STREAM": "0"}),
                     ("on", {"CRIS_B200_WGRAD_STREAM": "1"}), ("on2", {"CRIS_B200_WGRAD_STREAM": "1"}),
                     ("on_noprio", {"CRIS_B200_WGRAD_STREAM": "1", "CRIS_B200_WGRAD_PRIO": "0"})):
        os.environ.update(env)
        eng.graphs = {}
        for rep in range(3):
            with torch.no_grad():
                for k, b in model.named_buffers():
                    b.copy_(saved[k])
            model.zero_grad()
            pred, m, loss = model(img, word, mask)
            (loss * 3.0).backward()
            torch.cuda.synchronize()
            res[(tag, rep)] = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        os.environ.pop("CRIS_B200_WGRAD_PRIO", None)
    base = res[("off", 2)]
    for key in res:
        worst = sorted(((rel(res[key][k], base[k]), k) for k in base), reverse=True)[:4]
        print(key, " | ".join(f"{k.split('.')[-2][-14:]}.{k.split('.')[-1][:6]} {v:.2e}" for v, k in worst))
    k = "backbone.positional_embedding"
    for key in res:
        print(key, k, float(res[key][k].norm()), res[key][k].flatten()[:4].tolist())


if __name__ == "__main__":
    main()
