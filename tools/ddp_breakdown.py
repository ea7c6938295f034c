"""Per-phase timing of the SyncBN + DDP train step under torchrun (one rank per GPU): forward graph, backward
(graph + DDP all-reduce), optimizer, with and without DistributedDataParallel, to see where the N>1 overhead sits.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_breakdown.py
"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from oracle import synth  # noqa: E402


def run(mode, dev, rank, world, B=64):
    cfg, model, groups = bench.build_model("r50", dropout=0.1)
    model = model.to(dev)
    eng = model._get_engine()
    if mode != "nosync":
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    net = model
    if mode in ("ddp", "ddp_static"):
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=(mode == "ddp"),
                                                        static_graph=(mode == "ddp_static"), gradient_as_bucket_view=True)
    from cris.pytorch_b200.optim import Adam
    opt = Adam(groups, lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    img, word, mask = synth.make_inputs(B, rank, 416, cfg.word_len, synth.ARCHS["r50"]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)
    net.train()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = [0.0, 0.0, 0.0]
    n = 0
    for it in range(9):
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        pred, tgt, loss = net(img, word, mask)
        ev[1].record()
        opt.zero_grad()
        scaler.scale(loss).backward()
        ev[2].record()
        scaler.step(opt)
        scaler.update()
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 4:
            for k in range(3):
                acc[k] += ev[k].elapsed_time(ev[k + 1])
            n += 1
    t = torch.tensor([a / n for a in acc], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del net, model, opt
    eng.graphs = {}
    torch.cuda.empty_cache()
    return [float(x) for x in t]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for mode, what in (("nosync", "no SyncBN, no DDP (independent replicas)"), ("sync", "SyncBN exchange, no DDP"),
                       ("ddp", "SyncBN + DDP(find_unused_parameters=True)  [train.py:97-102]"),
                       ("ddp_static", "SyncBN + DDP(static_graph=True)")):
        try:
            f, b, o = run(mode, dev, rank, world)
            if rank == 0:
                print(f"{what:62s} fwd {f:6.2f}  bwd {b:6.2f}  opt {o:5.2f}  total {f + b + o:6.2f} ms", flush=True)
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"{what}: FAILED {e!r}"[:300], flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
