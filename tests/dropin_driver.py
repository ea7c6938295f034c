"""Runs the reference's OWN caller code against the drop-in `model` package (shim/model) — executed by
tests/test_dropin_gpu.py in a subprocess with PYTHONPATH = shim : repo : baseline/_ref : tests/stubs.

  * config: the reference's config/refcoco/cris_<arch>.yaml through the reference's utils/config.py
  * model:  `from model import build_segmenter`  (resolves to shim/model -> cris.pytorch_b200)
  * train:  two iterations of the reference's engine.engine.train (engine/engine.py:17-88) UNCHANGED, with the
            train.py:97-111 wrapping (SyncBatchNorm conversion, DistributedDataParallel, Adam, MultiStepLR, GradScaler)
  * eval :  the tools/latency.py:51-66 loop (shortened) on the same model
Prints one JSON line.
"""
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "r50"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ref_root = os.environ["CRIS_REF_ROOT"]
    import utils.config as config                       # the reference's loader
    from engine.engine import train                     # the reference's training loop, unchanged
    from model import build_segmenter                   # the drop-in
    import model as model_pkg
    from torch.optim.lr_scheduler import MultiStepLR
    from oracle import synth
    assert "shim" in model_pkg.__file__, model_pkg.__file__
    assert config.__file__.startswith(ref_root), config.__file__

    args = config.load_cfg_from_cfg_file(os.path.join(ref_root, "config", "refcoco", f"cris_{arch}.yaml"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(33500 + os.getpid() % 2000))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)   # train.py:80-83 (one process per GPU)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        args = config.merge_cfg_from_list(args, ["TRAIN.clip_pretrain", path, "TRAIN.print_freq", "1000000"])
        model, param_list = build_segmenter(args)            # train.py:94
    model.load_state_dict(synth.full_state_dict(arch, 0, synth.make_cfg(arch)), strict=True)
    if args.sync_bn:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)   # train.py:95-96
    model = torch.nn.parallel.DistributedDataParallel(model.cuda(), device_ids=[0],
                                                      find_unused_parameters=True)  # train.py:99-101
    optimizer = torch.optim.Adam(param_list, lr=args.base_lr, weight_decay=args.weight_decay)   # train.py:104-106
    scheduler = MultiStepLR(optimizer, milestones=args.milestones, gamma=args.lr_decay)         # train.py:107-109
    scaler = torch.amp.GradScaler("cuda")                                                        # train.py:110
    img, word, mask = synth.make_inputs(batch, 0, args.input_size, args.word_len, synth.ARCHS[arch]["vocab"])
    loader = [(img, word, mask.squeeze(1)) for _ in range(2)]   # RefDataset yields (img, word_vec, mask[H,W])
    before = {k: v.detach().clone() for k, v in list(model.module.named_parameters())[:5]
              if k != "backbone.logit_scale"}   # never used by the forward: no gradient, no update
    args.epochs = 1
    t0 = time.time()
    train(loader, model, optimizer, scheduler, scaler, 1, args)   # engine/engine.py:17-88, two iterations
    torch.cuda.synchronize()
    changed = sum(int(not torch.equal(before[k], dict(model.module.named_parameters())[k].detach())) for k in before)
    out = {"arch": arch, "train_iters": 2, "train_seconds": time.time() - t0, "params_changed": changed,
           "loss_scale": float(scaler.get_scale())}
    # tools/latency.py:38-66 (shortened): eval, no_grad, synchronize per call
    net = model.module.eval()
    image = torch.randn(1, 3, 416, 416).cuda()
    text = torch.randint(4096, size=(1, args.word_len)).long().cuda()
    ts = []
    with torch.no_grad():
        for i in range(60):
            s = time.time()
            _ = net(image, text)
            torch.cuda.synchronize()
            if i >= 10:
                ts.append(time.time() - s)
    out["latency_b1_ms"] = 1e3 * sorted(ts)[len(ts) // 2]
    out["pred_shape"] = list(_.shape)
    out["n_params_M"] = sum(p.numel() for p in net.parameters() if p.requires_grad) * 1e-6   # latency.py:34-35
    print("DROPIN " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
