"""Inference latency of the hot path (BASELINE.json config #5; mirrors the reference's tools/latency.py:52-75:
500 eval forwards of a 416x416 image + one sentence, the first 100 discarded, wall clock with a device
synchronize per call).  Prints p50 / mean latency and FPS for each requested batch size as one JSON line.

  python tools/latency.py [--arch r50] [--batches 1 32] [--iters 500]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="r50")
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 32])
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--eager", action="store_true", help="disable the captured inference graph")
    args = ap.parse_args()
    import bench
    from oracle import synth
    torch.cuda.set_device(0)
    cfg, model, _ = bench.build_model(args.arch, dropout=0.0)
    model = model.cuda().eval()
    if args.eager:
        model._get_engine().use_graphs = False
    size = 128 if args.arch == "tiny" else 416
    out = {"arch": args.arch, "graphs": not args.eager, "results": []}
    for b in args.batches:
        image = torch.randn(b, 3, size, size).cuda()
        text = torch.randint(1, synth.ARCHS[args.arch]["vocab"] - 2, size=(b, cfg.word_len)).long().cuda()
        ts = []
        with torch.no_grad():
            for i in range(args.iters):
                t0 = time.perf_counter()
                _ = model(image, text)
                torch.cuda.synchronize()
                if i >= args.iters // 5:
                    ts.append(time.perf_counter() - t0)
        ts.sort()
        out["results"].append({"batch": b, "p50_ms": 1e3 * ts[len(ts) // 2], "mean_ms": 1e3 * sum(ts) / len(ts),
                               "p99_ms": 1e3 * ts[int(len(ts) * 0.99)], "fps": b * len(ts) / sum(ts)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
