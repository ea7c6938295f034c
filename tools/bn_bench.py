"""Micro-benchmark of the three BatchNorm passes through the C ABI at the model's shapes (B=64): achieved HBM GB/s
against MEASURED_PEAKS.json.  Algorithmic bytes: apply = z (+resid) read + y write; reduce = dy + z (+y) read;
backward apply = dy + z (+y) read + dz (+dres) write.   python tools/bn_bench.py [--stream 0|1]"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cris.pytorch_b200._lib import call  # noqa: E402

SHAPES = [(64, 208, 208, 32), (64, 104, 104, 64), (64, 104, 104, 256), (64, 52, 52, 512), (64, 26, 26, 1024), (64, 13, 13, 2048),
          (64, 104, 104, 128)]


def timed(fn, iters=20):
    if "--once" in sys.argv:  # under ncu: one launch per variant
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    if "--stream" in sys.argv:
        os.environ["CRIS_B200_BN_STREAM"] = sys.argv[sys.argv.index("--stream") + 1]
    try:
        peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    out = []
    for N, H, W, C in (SHAPES[2:3] if "--once" in sys.argv else SHAPES):
        hp, wp = H + 2, W + 2
        rows = N * hp * wp
        mk = lambda: torch.randn(rows, C, device="cuda").to(torch.bfloat16)  # noqa: E731
        z, resid, dy, y, dz, dres = mk(), mk(), mk(), mk(), mk(), mk()
        scale, shift, mean, invstd = (torch.rand(C, device="cuda") + 0.5 for _ in range(4))
        nb = max(1, min(592, rows // 64))
        part = torch.zeros(min(nb, 64) * 2 * C, device="cuda")
        sums = torch.randn(2 * C, device="cuda")
        tensor_gb = rows * C * 2 / 1e9
        for with_res in (False, True):
            t = timed(lambda: call("cris_bn_apply", z.data_ptr(), C, scale.data_ptr(), shift.data_ptr(),
                                   resid.data_ptr() if with_res else None, C, y.data_ptr(), C, rows, C, 1, hp, wp))
            gb = tensor_gb * (3 if with_res else 2)
            out.append(("apply" + ("+resid" if with_res else ""), N, H, W, C, t * 1e6, gb / t, gb / t / peak))
            ym = y if with_res else None
            t = timed(lambda: call("cris_col_reduce", 1, dy.data_ptr(), C, 0, None, 0, ym.data_ptr() if ym is not None else None,
                                   C if ym is not None else 0, z.data_ptr(), C, 0, mean.data_ptr(), invstd.data_ptr(),
                                   scale.data_ptr(), shift.data_ptr(), rows, C, 1, hp, wp, part.data_ptr(), nb))
            gb = tensor_gb * (3 if with_res else 2)
            out.append(("reduce" + ("+y" if with_res else ""), N, H, W, C, t * 1e6, gb / t, gb / t / peak))
            t = timed(lambda: call("cris_bn_bwd_apply", dy.data_ptr(), C, ym.data_ptr() if ym is not None else None,
                                   C if ym is not None else 0, z.data_ptr(), C, mean.data_ptr(), invstd.data_ptr(),
                                   scale.data_ptr(), shift.data_ptr(), sums.data_ptr(), float(N * H * W), dz.data_ptr(), C,
                                   dres.data_ptr() if with_res else None, C, 0, rows, C, 1, hp, wp))
            gb = tensor_gb * (5 if with_res else 3)
            out.append(("bwd_apply" + ("+y+dres" if with_res else ""), N, H, W, C, t * 1e6, gb / t, gb / t / peak))
        del z, resid, dy, y, dz, dres
    print(f"BN_STREAM={os.environ.get('CRIS_B200_BN_STREAM', '1')}  HBM peak {peak:.0f} GB/s")
    for r in out:
        print(f"{r[0]:18s} {r[1]}x{r[2]}x{r[3]}x{r[4]:<5d} {r[5]:8.1f} us  {r[6]:7.0f} GB/s  {r[7]:.2f} of peak")


if __name__ == "__main__":
    main()
