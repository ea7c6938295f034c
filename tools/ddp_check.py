"""Multi-GPU consistency check (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_check.py

Trains the same SyncBN+DDP model (train.py:97-102 recipe, dropout 0 so the runs are comparable) for a few steps
twice from identical weights: (a) eager launches with torch.distributed/NCCL statistics exchange, (b) captured CUDA
graphs with the NVLink peer-memory exchange (csrc/peer.cu).  Both must give the same losses (fp32 summation order
of the statistics differs: NCCL ring vs rank order) and identical parameters on every rank.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(mode, arch, B, size, steps, dev, rank, world):
    import bench
    from oracle import synth
    os.environ["CRIS_B200_PEER"] = "1" if mode == "peer+graphs" else "0"
    torch.manual_seed(0)
    cfg, model, groups = bench.build_model(arch, dropout=0.0)
    model = model.to(dev)
    eng = model._get_engine()
    eng.use_graphs = mode == "peer+graphs"
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)
    opt = torch.optim.Adam(groups, lr=1e-4)
    img, word, mask = synth.make_inputs(B, rank, size, cfg.word_len, synth.ARCHS[arch]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)
    model.train()
    losses = []
    for _ in range(steps):
        pred, tgt, loss = model(img, word, mask)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ld = loss.detach().clone()
        dist.all_reduce(ld)
        losses.append(ld.item() / world)
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
    digest = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    graphed = bool(eng.graphs)
    del model, opt
    eng.graphs = {}
    torch.cuda.empty_cache()
    return losses, same, graphed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="r50")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=0, help="default: the arch's native input size")
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    args.size = args.size or (128 if args.arch == "tiny" else 416)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    a = run("nccl+eager", args.arch, args.batch, args.size, args.steps, dev, rank, world)
    b = run("peer+graphs", args.arch, args.batch, args.size, args.steps, dev, rank, world)
    rel = max(abs(x - y) / max(abs(x), 1e-6) for x, y in zip(a[0], b[0]))
    ok = a[1] and b[1] and b[2] and not a[2] and rel < 2e-2
    if rank == 0:
        print(json.dumps({"world": world, "nccl_eager_losses": a[0], "peer_graph_losses": b[0], "max_rel_diff": rel,
                          "params_identical_across_ranks": [a[1], b[1]], "graphs_used": [a[2], b[2]], "ok": ok}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
