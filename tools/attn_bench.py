"""Micro-benchmark of the fused attention kernels (csrc/attention.cu) at the decoder self-attention shape
(B=64, 8 heads, 676 x 676, d=64, dropout 0.1) and the attention-pool shape (B=64, 32 heads, 169 x 169): ms, and
TFLOP/s over the algorithmic flop (fwd 4*Lq*Lk*64 per head incl. none of the recompute; bwd 10*Lq*Lk*64).
   python tools/attn_bench.py [--once]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cris.pytorch_b200._lib import call  # noqa: E402


def timed(fn, iters=10):
    if "--once" in sys.argv:
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    shapes = [(64, 8, 676, 676, 0.1)] if "--once" in sys.argv else [(64, 8, 676, 676, 0.1), (64, 8, 676, 676, 0.0), (64, 32, 169, 169, 0.0)]
    for B, H, Lq, Lk, pd in shapes:
        E = H * 64
        mk = lambda rows: torch.randn(rows, E, device="cuda").to(torch.bfloat16)  # noqa: E731
        q, k, v, do = mk(B * Lq), mk(B * Lk), mk(B * Lk), mk(B * Lq)
        o = torch.empty_like(q)
        lse = torch.empty(B * H * Lq, device="cuda")
        dsum = torch.empty(B * H * Lq, device="cuda")
        dq = torch.zeros(B * Lq, E, device="cuda")
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        seed = 12345
        tf = timed(lambda: call("cris_attention_fwd", q.data_ptr(), E, k.data_ptr(), E, v.data_ptr(), E, o.data_ptr(), E,
                                lse.data_ptr(), B, H, Lq, Lk, 0.125, pd, seed, None))
        tb = timed(lambda: call("cris_attention_bwd", q.data_ptr(), E, k.data_ptr(), E, v.data_ptr(), E, o.data_ptr(), E,
                                do.data_ptr(), E, lse.data_ptr(), dsum.data_ptr(), dq.data_ptr(), E, dk.data_ptr(), E,
                                dv.data_ptr(), E, B, H, Lq, Lk, 0.125, pd, seed, None))
        fl = B * H * Lq * Lk * 64 * 2.0
        print(f"B={B} heads={H} Lq={Lq} Lk={Lk} p_drop={pd}: fwd {tf:.3f} ms ({2 * fl / tf * 1e-9:.0f} TFLOP/s)  "
              f"bwd {tb:.3f} ms ({5 * fl / tb * 1e-9:.0f} TFLOP/s)")


if __name__ == "__main__":
    main()
