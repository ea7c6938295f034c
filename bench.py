#!/usr/bin/env python
"""bench.py — CRIS train-step throughput on B200 (BASELINE.json metric: images/sec, 416x416, len 17, bs 64/GPU).

    python bench.py --gpus N --steps K --warmup W                   # product arm (cris.pytorch_b200)
    python bench.py --impl reference --gpus N --steps K --warmup W  # reference arm: the UNMODIFIED reference
                                                                    # (baseline/_ref) on the box's host cores
    python bench.py --impl incumbent [--arch r50] [--batch 64]      # the unmodified reference, eager PyTorch on
                                                                    # the same GPU (cuDNN/cuBLAS, autocast) — the
                                                                    # kernel library path this build has to beat

One "step" = the reference's training iteration (engine/engine.py:48-70): forward, optimizer.zero_grad,
scaler.scale(loss).backward(), scaler.step(Adam), scaler.update(), trainMetricGPU, 3 scalar all-reduces and the
3 .item() host reads — on synthetic RefCOCO-shaped data (there is no network for datasets / checkpoints).
Prints ONE JSON line (rank 0).  `value` = steps with inputs resident in HBM; `e2e` = the same step fed from
pinned HOST buffers (H2D copies + D2H scalar reads inside the timed region).

`roofline` is the WHOLE STEP against the sustained bf16 tensor peak (MEASURED_PEAKS.json): images/s x F_train.
Under it: `dominant_kernel` (the tcgen05 implicit GEMM of proj.vis.3, the largest convolution, timed live with CUDA
events on its launch stream inside full steps) and `classes` — every launch of one eagerly-launched step timed with
its own event pair and summed per class (GEMM fwd / dgrad / wgrad / attention products, BatchNorm, LayerNorm,
softmax, ...; GEMM classes carry their algorithmic flop -> TFLOP/s and fraction of peak).
`gpu_incumbent` = the unmodified reference run eagerly on the same GPU in the same bench invocation (N=1).
`cpu_baseline` = the unmodified reference's train step on the box's host cores (bounded sample).
`extra` = inference latency (tools/latency.py recipe) at batch 1 and 32 with its roofline fraction.
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
F_FWD_GF = {"r50": 131.9, "r101": 156.64}
METRIC = "images/sec (train step, 416x416, len=17, bs=64/GPU)"

KERNEL_CLASS = {  # C-ABI entry point -> class of the per-class table
    "cris_col_reduce": "batchnorm/bias reductions", "cris_bn_finalize_fwd": "batchnorm", "cris_stats_finalize_bwd": "batchnorm",
    "cris_bn_reduce_partials": "batchnorm", "cris_bn_coeffs": "batchnorm", "cris_bn_apply": "batchnorm",
    "cris_bn_bwd_apply": "batchnorm", "cris_bn_bwd_fused": "batchnorm", "cris_layernorm_fwd": "layernorm",
    "cris_layernorm_bwd": "layernorm", "cris_softmax_fwd": "softmax", "cris_softmax_bwd": "softmax",
    "cris_attention_fwd": "fused attention", "cris_attention_bwd": "fused attention",
    "cris_avgpool2_fwd": "resample", "cris_avgpool2_bwd": "resample", "cris_upsample2x_fwd": "resample",
    "cris_upsample2x_bwd": "resample", "cris_elementwise": "elementwise", "cris_pack_conv_weight": "weight repack",
    "cris_pack_matrix": "weight repack", "cris_dynconv_bce_fwd": "dynconv+BCE", "cris_dynconv_bce_bwd": "dynconv+BCE",
    "cris_conv3x3_halo": "gemm_fwd", "cris_peer_allreduce_f32": "syncbn exchange",
}


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


# ---------------------------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (baseline/_ref) when it is staged, else the oracle port
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_steps(arch, batch, steps, warmup, threads, size=416):
    """Fallback only (baseline/_ref absent): the oracle port of the step (fp32, torch CPU)."""
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


def cpu_reference(arch, steps, warmup, batch=2):
    """-> (images/s, cpu_baseline dict).  The reference's train step on the host cores, bounded sample."""
    from baseline import ref_step
    from oracle.hostinfo import cpu_model, usable_cpus
    cores = usable_cpus()
    size = 128 if arch == "tiny" else 416
    if ref_step.available():
        r = ref_step.train_steps(arch, batch, steps, warmup, "cpu", None, cores, size)
        secs, kind, what = r["seconds"], "reference", "the unmodified reference (baseline/_ref, model.build_segmenter + engine.py:48-70 step)"
    else:
        secs, _ = cpu_oracle_steps(arch, batch, steps, warmup, cores, size)
        kind, what = "port", "oracle port of model/segmenter.py (baseline/_ref not staged)"
    v = batch * steps / secs
    sample = (f"{steps} train step(s) (fwd+loss+bwd+Adam+metric, fp32 torch CPU) of {what} at batch {batch} "
              f"(BatchNorm1d needs >= 2), {size}x{size}, len 17, on {cores} threads of {cpu_model()}")
    return v, secs, {"value": v, "unit": "images/sec", "cores": cores, "kind": kind, "sample": sample}


def run_reference(args, rank):
    if rank != 0:
        return
    v, secs, cb = cpu_reference(args.arch, args.steps, min(args.warmup, 1))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cris_{args.arch} train step (fwd+BCE loss+bwd+Adam+metrics), 416x416, word_len 17, "
                               "bounded CPU sample at batch 2 of the batch-64/GPU workload", "global_batch": 2,
                   "parallelism": "cpu"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_incumbent(args):
    """The unmodified reference, eager PyTorch on cuda:0: stock fp16 autocast + GradScaler (train.py:111,
    engine.py:48-57) and bf16 autocast, cudnn.benchmark on, the same synthetic weights / batch / step."""
    from baseline import ref_step
    if not ref_step.available():
        print(json.dumps({"impl": "incumbent", "unavailable": "baseline/_ref is not staged"}))
        return
    torch.cuda.set_device(0)
    B = args.batch or (64 if args.arch != "r101" else 32)
    out = {"impl": "incumbent", "arch": args.arch, "batch": B, "steps": args.steps, "warmup": args.warmup,
           "what": "unmodified reference (baseline/_ref) through model.build_segmenter, eager PyTorch "
                   f"{torch.__version__} on the same GPU: cuDNN/cuBLAS kernels, torch.autocast, cudnn.benchmark=True, "
                   "torch.optim.Adam, GradScaler (fp16), engine/engine.py:48-70 step"}
    for name, dt in (("fp16_stock", torch.float16), ("bf16", torch.bfloat16)):
        try:
            r = ref_step.train_steps(args.arch, B, args.steps, args.warmup, "cuda", dt)
            out[name] = {"images_per_sec": B * args.steps / r["seconds"], "ms_per_step": 1e3 * r["seconds"] / args.steps,
                         "loss_last": r["losses"][-1]}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    if args.latency:
        lat = {}
        for b in (1, 32):
            try:
                lat[f"b{b}_p50_ms_fp16"] = ref_step.eval_latency(args.arch, b, 150, "cuda", torch.float16)
            except Exception as e:  # noqa: BLE001
                lat[f"b{b}_error"] = repr(e)[:200]
        out["latency"] = lat
    best = max((v["images_per_sec"] for k, v in out.items() if isinstance(v, dict) and "images_per_sec" in v), default=None)
    out["images_per_sec"] = best
    print("INCUMBENT " + json.dumps(out), flush=True)


def spawn_incumbent(arch, batch, latency):
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "incumbent", "--arch", arch, "--batch", str(batch),
           "--steps", "5", "--warmup", "3"] + (["--latency"] if latency else [])
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = next((ln for ln in r.stdout.splitlines() if ln.startswith("INCUMBENT ")), None)
        if line is None:
            return {"error": (r.stderr or r.stdout)[-400:]}
        return json.loads(line[len("INCUMBENT "):])
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


# ---------------------------------------------------------------------------------------------------------------
def class_table(rows, pk):
    """rows: [(entry point, class, flop, ms)] of ONE eagerly launched step -> per-class summary."""
    agg = {}
    for name, cls, flop, ms in rows:
        c = cls or KERNEL_CLASS.get(name, "other")
        a = agg.setdefault(c, {"ms": 0.0, "launches": 0, "flop": 0.0})
        a["ms"] += ms
        a["launches"] += 1
        a["flop"] += flop
    total = sum(a["ms"] for a in agg.values()) or 1.0
    out = {}
    for c, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(a["ms"], 3), "share": round(a["ms"] / total, 4), "launches": a["launches"]}
        if a["flop"] > 0:
            tf = a["flop"] / (a["ms"] * 1e-3) / 1e12
            e["tflops"] = round(tf, 1)
            e["frac_of_peak"] = round(tf / pk["bf16_tflops"], 4)
        out[c] = e
    return out, total


def latency_extra(arch, model, cfg, pk):
    """tools/latency.py:51-66 recipe (eval, no_grad, sync per call) at batch 1 and 32, p50 of the last 80 %."""
    from oracle import synth
    model.eval()
    size = 128 if arch == "tiny" else 416
    res = {}
    for b, iters in ((1, 200), (32, 100)):
        image = torch.randn(b, 3, size, size).cuda()
        text = torch.randint(1, synth.ARCHS[arch]["vocab"] - 2, size=(b, cfg.word_len)).long().cuda()
        ts = []
        with torch.no_grad():
            for i in range(iters):
                t0 = time.perf_counter()
                model(image, text)
                torch.cuda.synchronize()
                if i >= iters // 5:
                    ts.append(time.perf_counter() - t0)
        ts.sort()
        p50 = ts[len(ts) // 2]
        tf = b * F_FWD_GF.get(arch, 0.0) / 1e3 / p50
        res[f"b{b}"] = {"p50_ms": round(1e3 * p50, 3), "images_per_sec": round(b / p50, 1), "tflops": round(tf, 1),
                        "frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4)}
    model.train()
    return res


def r101_extra(dev, pk, B=32, steps=5, warmup=3):
    """cris_r101 train step at the reference config's 32 images per GPU (config/refcoco/cris_r101.yaml batch_size 32 x 8)."""
    from oracle import synth
    from cris.pytorch_b200.optim import Adam
    cfg, model, groups = build_model("r101", dropout=0.1)
    model = model.to(dev).train()
    opt = Adam(groups, lr=1e-4, weight_decay=0.0)
    scaler = torch.amp.GradScaler("cuda")
    img, word, mask = synth.make_inputs(B, 0, 416, cfg.word_len, synth.ARCHS["r101"]["vocab"])
    img, word, mask = img.to(dev), word.to(dev), mask.to(dev)

    def step():
        pred, tgt, loss = model(img, word, mask)
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        model.train_metric()
        return loss.item()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    v = B / (ms / 1e3)
    tf = v * F_TRAIN_GF["r101"] / 1e3
    res = {"images_per_sec": round(v, 1), "ms_per_step": round(ms, 2), "batch": B, "tflops": round(tf, 1),
           "frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4),
           "workload": "cris_r101 train step (fwd+BCE loss+bwd+Adam+metric), 416x416, word_len 17, batch 32/GPU (BASELINE.json config 4, per GPU)"}
    model._get_engine().graphs = {}
    del model, opt
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "incumbent"])
    ap.add_argument("--arch", default="r50", choices=["r50", "r101", "tiny"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 64; r101 config: 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-incumbent", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-class table and the inference latency legs")
    ap.add_argument("--latency", action="store_true", help="(incumbent) also time the tools/latency.py recipe")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "incumbent":
        run_incumbent(args)
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
    bare = model
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
        iou, pr5 = bare.train_metric()                        # :60 trainMetricGPU, counts fused into the loss kernel
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
    # the fused metric is the reference's trainMetricGPU: check it once against the torch restatement of utils/misc.py
    with torch.no_grad():
        p_chk, t_chk, _ = model(img_d, word_d, mask_d)
        a, b_ = bare.train_metric()
        c, d = train_metric(p_chk, t_chk)
        assert abs(float(a) - float(c)) < 0.05 and abs(float(b_) - float(d)) < 2.0, (float(a), float(c), float(b_), float(d))
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
    pk, pk_src = peaks()
    # every launch of one eager step with its own event pair -> per-class table (all ranks run it: SyncBN sites)
    classes, classes_total = None, None
    if not args.no_extras:
        step(img_d, word_d, mask_d)
        _lib.profile_begin()
        step(img_d, word_d, mask_d)
        rows = _lib.profile_end()
        classes, classes_total = class_table(rows, pk)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
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
        "roofline": {"bound": "tensor", "kernel": "WHOLE STEP (every launch of fwd+bwd+Adam): images/s/GPU x F_train "
                                                    f"({F_TRAIN_GF.get(args.arch)} GF/img) vs the sustained bf16 peak",
                     "achieved": step_tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": step_tflops / pk["bf16_tflops_sustained"], "traffic": None,
                     "peak_source": pk_src + " (sustained: the kernels run inside a long step)",
                     "dominant_kernel": {
                         "kernel": "gemm_tc_kernel<256,64,K-major,K-major> implicit-GEMM conv3x3 proj.vis.3 fwd (persistent tcgen05)",
                         "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                         "traffic": traffic, "kernel_ms": k_ms, "flop_per_launch": k_flop,
                         "share_of_step": k_ms / (ms_total / args.steps) if ms_total > 0 else None,
                         "peak_source": pk_src + " (burst: kernel timed alone by events)"},
                     "classes": classes, "classes_total_ms": classes_total,
                     "classes_note": "one eagerly launched step, one CUDA-event pair per launch (sum of kernel times; "
                                     "the graph-replayed step above is what `value` times)"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_extras:
        try:
            out["extra"] = {"inference_latency": latency_extra(args.arch, bare, cfg, pk),
                            "recipe": "tools/latency.py:51-66 of the reference (eval, no_grad, synchronize per call, p50)"}
        except Exception as e:  # noqa: BLE001
            out["extra"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_extras and args.arch == "r50":
        # BASELINE.json config #4 (cris_r101, 32 images per GPU) on this GPU: the same step, 3 warm-up + 5 timed
        try:
            engine.graphs, engine.eval_graphs = {}, {}
            del model, opt
            torch.cuda.empty_cache()
            out["extra"]["r101"] = r101_extra(dev, pk)
        except Exception as e:  # noqa: BLE001
            out.setdefault("extra", {})["r101"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_incumbent:
        # free this arm's captured graphs first: the incumbent needs its own ~60 GB of eager activations
        engine.graphs, engine.eval_graphs = {}, {}
        torch.cuda.empty_cache()
        inc = spawn_incumbent(args.arch, B, latency=not args.no_extras)
        out["gpu_incumbent"] = inc
        if inc.get("images_per_sec"):
            out["gpu_incumbent"]["speedup_vs_incumbent"] = value / inc["images_per_sec"]
    if not args.no_cpu_baseline and world == 1:
        try:
            _, _, out["cpu_baseline"] = cpu_reference(args.arch, 2, 1)
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"error": repr(e)[:300]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
