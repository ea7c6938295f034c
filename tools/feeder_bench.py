"""Time the batched GPU input transform (cris.pytorch_b200.feeder.letterbox_batch) on 64 synthetic 480x640 photos + masks,
next to the reference's per-sample cv2 + torch lines (utils/dataset.py:148-163,210-221) on the host.
    python tools/feeder_bench.py [batch]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cris.pytorch_b200 import _lib  # noqa: E402
from cris.pytorch_b200.feeder import BORDER, MEAN, STD, letterbox_batch  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    S = 416
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(B)]
    masks = [(rng.integers(0, 2, (480, 640), dtype=np.uint8) * 255) for _ in range(B)]
    sc = min(S / 480, S / 640)
    mat = np.array([[sc, 0, (S - 640 * sc) / 2], [0, sc, (S - 480 * sc) / 2]], np.float64)
    mats = [mat] * B
    for _ in range(3):
        out, om = letterbox_batch(imgs, mats, masks, input_size=S)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out, om = letterbox_batch(imgs, mats, masks, input_size=S)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / n * 1e3
    # kernel alone: profile hooks of the ctypes layer (one CUDA-event pair around the launch)
    _lib.profile_begin()
    letterbox_batch(imgs, mats, masks, input_size=S)
    rows = _lib.profile_end()
    kern = sum(r[3] for r in rows if r[0] == "cris_feeder_letterbox")
    in_mb = sum(i.nbytes + m.nbytes for i, m in zip(imgs, masks)) / 1e6
    out_mb = (out.numel() + om.numel()) * 4 / 1e6
    print(f"GPU letterbox_batch B={B}: kernel {kern:.3f} ms ({(in_mb + out_mb) / kern:.0f} GB/s of uint8-in + fp32-out), "
          f"whole call incl. packing + H2D of {in_mb:.0f} MB: {e2e:.1f} ms ({B / e2e * 1e3:.0f} samples/s)")
    try:
        import cv2
    except ImportError:
        print("cv2 not installed: no host comparison")
        return
    mean = torch.tensor(MEAN).reshape(3, 1, 1)
    std = torch.tensor(STD).reshape(3, 1, 1)
    cv2.setNumThreads(1)   # one loader worker = one core in the reference's DataLoader
    t0 = time.perf_counter()
    for b in range(B):
        wi = cv2.warpAffine(imgs[b], mat, (S, S), flags=cv2.INTER_CUBIC, borderValue=list(BORDER))
        wm = cv2.warpAffine(masks[b], mat, (S, S), flags=cv2.INTER_LINEAR, borderValue=0.) / 255.
        t = torch.from_numpy(wi.transpose((2, 0, 1))).float()
        t.div_(255.).sub_(mean).div_(std)
        torch.from_numpy(wm).float()
    host = (time.perf_counter() - t0) * 1e3
    print(f"host cv2 + torch, one thread: {host:.1f} ms per {B} samples ({B / host * 1e3:.0f} samples/s per loader worker); "
          f"fp32 tensors to ship instead: {out_mb:.0f} MB")


if __name__ == "__main__":
    main()
