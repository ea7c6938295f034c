#!/usr/bin/env python
"""bench.py — CRIS train-step throughput on B200 (BASELINE.json metric: images/sec, 416x416, len 17, bs 64/GPU).

    python bench.py --gpus N --steps K --warmup W                 # product arm (cris.pytorch_b200)
    python bench.py --impl reference --gpus N --steps K --warmup W  # reference arm: the CPU oracle port

One "step" = the reference's training iteration (engine/engine.py:48-70): forward, optimizer.zero_grad,
scaler.scale(loss).backward(), scaler.step(Adam), scaler.update(), trainMetricGPU, 3 scalar all-reduces and the
3 .item() host reads — on synthetic RefCOCO-shaped data (there is no network for datasets / checkpoints).
Prints ONE JSON line (rank 0).  `value` = steps with inputs resident in HBM; `e2e` = the same step fed from
pinned HOST buffers (H2D copies + D2H scalar reads inside the timed region).  `roofline` = the dominant
kernel (the tcgen05 implicit-GEMM of proj.vis.3, the largest convolution) timed live with CUDA events inside
the timed steps; `cpu_baseline` = the oracle port of the same step on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

F_TRAIN_GF = {"r50": 395.7, "r101": 469.85}  # GFLOP per image fwd+bwd (SURVEY.md §8d / BASELINE.md §2)
METRIC = "images/sec (train step, 416x416, len=17, bs=64/GPU)"


def peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def train_metric(pred, target, threshold=0.35, pr_iou=0.5):
    """utils/misc.py:114-129 trainMetricGPU restated (sigmoid, threshold, per-sample IoU, Pr@50)."""
    o = (torch.sigmoid(pred.flatten(1)) >= threshold)
    t = target.flatten(1).bool()
    inter = (o & t).sum(1)
    union = (o | t).sum(1)
    ious = inter / (union + 1e-6)
    return 100.0 * ious.mean(), 100.0 * (ious > pr_iou).float().mean()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        load = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(arch, dropout):
    from oracle import synth  # synthetic weights only (no checkpoints on a no-network box)
    from cris.pytorch_b200 import build_segmenter
    cfg = synth.make_cfg(arch, dropout=dropout)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.pt")
        synth.save_clip_torchscript(synth.clip_state_dict(arch, 0), path)
        cfg.clip_pretrain = path
        model, groups = build_segmenter(cfg)
    model.load_state_dict(synth.full_state_dict(arch, 0, cfg), strict=True)
    return cfg, model, groups


def cpu_reference_steps(arch, batch, steps, warmup, threads, size=416):
    """The reference's CPU implementation of the step = the oracle port (fp32, torch CPU), timed on the host."""
    from oracle import cris_oracle as O, synth
    torch.set_num_threads(threads)
    cfg = synth.make_cfg(arch, dropout=0.1)
    sd = synth.full_state_dict(arch, 0, cfg)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and "running" not in k and k != "backbone.logit_scale"}
    live = dict(sd)
    live.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    img, word, mask = synth.make_inputs(batch, 0, size, cfg.word_len, synth.ARCHS[arch]["vocab"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.cris_forward(live, img, word, mask, training=True, num_head=cfg.num_head, dropout_p=0.1)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        for k, v in out["new_running"].items():
            live[k] = v
        train_metric(out["pred"].detach(), out["mask"])
        float(out["loss"].detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times), batch


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle.hostinfo import cpu_model, usable_cpus
    cores = usable_cpus()
    batch = 2 if (args.steps + args.warmup) <= 8 else 1
    total, b = cpu_reference_steps(args.arch, batch, args.steps, args.warmup, cores, 416 if args.arch != "tiny" else 128)
    v = b * args.steps / total
    sample = (f"{args.steps} train steps (fwd+loss+bwd+Adam, fp32 torch CPU oracle port of model/segmenter.py) at batch {b}, "
              f"416x416, len 17 on {cores} threads of {cpu_model()}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cris_{args.arch} train step, 416x416, word_len 17, CPU sample batch {b}",
                   "global_batch": b, "parallelism": "cpu"},
        "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="r50", choices=["r50", "r101", "tiny"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 64; r101 config: 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    import torch.distributed as dist
    from cris.pytorch_b200 import _lib
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or (64 if args.arch != "r101" else 32)
    size = 416 if args.arch != "tiny" else 128
    from oracle import synth
    torch.manual_seed(rank)
    cfg, model, groups = build_model(args.arch, dropout=0.1)
    model = model.to(dev)
    engine = model._get_engine()
    if world > 1:  # train.py:97-102
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    if os.environ.get("CRIS_B200_TORCH_ADAM", "0") == "1":
        opt = torch.optim.Adam(groups, lr=1e-4, weight_decay=0.0)  # train.py:105-107, stock optimizer
    else:  # same update, one launch per parameter group (cris/pytorch_b200/optim.py)
        from cris.pytorch_b200.optim import Adam
        opt = Adam(groups, lr=1e-4, weight_decay=0.0)
    scaler = torch.amp.GradScaler("cuda")                          # train.py:111
    vocab = synth.ARCHS[args.arch]["vocab"]
    img_h, word_h, mask_h = synth.make_inputs(B, rank, size, cfg.word_len, vocab)
    img_h, word_h, mask_h = img_h.pin_memory(), word_h.pin_memory(), mask_h.squeeze(1).pin_memory()
    img_d, word_d, mask_d = img_h.to(dev), word_h.to(dev), mask_h.to(dev).unsqueeze(1)
    model.train()

    def step(image, text, target):
        pred, tgt, loss = model(image, text, target)          # engine.py:48-49
        opt.zero_grad()                                       # :52
        scaler.scale(loss).backward()                         # :53
        scaler.step(opt)                                      # :56
        scaler.update()                                       # :57
        iou, pr5 = train_metric(pred, tgt)                    # :60
        ld = loss.detach()
        if world > 1:                                         # :61-63
            dist.all_reduce(ld); dist.all_reduce(iou); dist.all_reduce(pr5)
        return (ld / world).item(), (iou / world).item(), (pr5 / world).item()   # :68-70

    def timed(n, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        if from_host:
            # engine.py:40-46 with the loader wrapped in DevicePrefetcher: every step's inputs are copied from pinned
            # host memory inside the timed region (n batches -> n copies of 177 MB), one batch ahead on a side stream
            from cris.pytorch_b200.data import DevicePrefetcher
            for image, text, target in DevicePrefetcher(((img_h, word_h, mask_h) for _ in range(n)), dev):
                image = image.cuda(non_blocking=True)          # no-ops on device tensors, as in the reference loop
                text = text.cuda(non_blocking=True)
                target = target.cuda(non_blocking=True).unsqueeze(1)
                last = step(image, text, target)
        else:
            for _ in range(n):
                last = step(img_d, word_d, mask_d)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms), last

    for _ in range(args.warmup):
        step(img_d, word_d, mask_d)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    ms_total, last = timed(args.steps, from_host=False)
    launches = _lib.launch_count() - l0
    timed(4, from_host=True)  # untimed warm-up of the host-fed path (staging buffers of the copy stream get allocated)
    ms_e2e, _ = timed(args.steps, from_host=True)
    # dominant kernel, timed live with CUDA events on its launch stream: the steps above replay CUDA graphs (no
    # per-kernel events possible inside a replay), so the same step is run eagerly here with an event pair around
    # the forward launch of the largest convolution
    probe = "proj.vis.3.0.weight"
    engine.probe_name, engine.probe_events = probe, []
    for _ in range(3):
        step(img_d, word_d, mask_d)
    torch.cuda.synchronize()
    engine.probe_name = None
    probe_ms = [a.elapsed_time(b) for a, b in engine.probe_events]
    engine.probe_events = []
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_src = peaks()
    value = world * B * args.steps / (ms_total / 1000.0)
    e2e = world * B * args.steps / (ms_e2e / 1000.0)
    # dominant kernel: forward implicit GEMM of proj.vis.3 (3x3, 2C->C at H/4 x W/4): 2*B*Ho*Wo*Cout*Cin*9 flop
    c = cfg.vis_dim // 2
    ho = size // 4
    k_flop = 2.0 * B * ho * ho * c * (2 * c) * 9
    k_ms = sum(probe_ms) / max(1, len(probe_ms))
    ach = k_flop / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "top_kernel_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    step_tflops = value * F_TRAIN_GF.get(args.arch, 0.0) / 1e3 / world
    out = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"cris_{args.arch} train step (fwd+BCE loss+bwd+Adam+metrics), {size}x{size}, word_len "
                               f"{cfg.word_len}, batch {B}/GPU, synthetic RefCOCO-shape data, dropout 0.1, SyncBN+DDP when N>1",
                   "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": "no flush needed: one step streams >20 GB of activations through a 126 MB L2",
                   "loss_last_step": last[0] if last else None},
        "e2e": {"value": e2e, "unit": "images/sec", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(img_h.numel() * 4 + word_h.numel() * 8 + mask_h.numel() * 4),
                "d2h_bytes_per_step": 12},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel<256,64,K-major,K-major> implicit-GEMM conv3x3 proj.vis.3 fwd (persistent tcgen05)",
                     "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                     "traffic": traffic, "peak_source": pk_src + " (burst, kernel timed alone by events)",
                     "kernel_ms": k_ms, "flop_per_launch": k_flop,
                     "step": {"achieved": step_tflops, "peak": pk["bf16_tflops_sustained"],
                              "frac": step_tflops / pk["bf16_tflops_sustained"],
                              "note": "whole step: images/s/GPU x F_train (395.7 GF/img r50) vs sustained bf16 peak"}},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.hostinfo import cpu_model, usable_cpus
        cores = usable_cpus()
        t, b = cpu_reference_steps(args.arch, 2, 1, 0, cores, size)
        out["cpu_baseline"] = {"value": b / t, "unit": "images/sec", "cores": cores, "kind": "port",
                               "sample": f"1 train step (fwd+loss+bwd+Adam) of cris_{args.arch} at batch {b}, {size}x{size}, "
                                         f"fp32 torch-CPU oracle port, {cores} threads of {cpu_model()}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
