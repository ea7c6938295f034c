"""GPU evaluation post-processing (the step right after the model in the reference's validate()/inference(),
engine/engine.py:101-124,172-190): sigmoid -> bicubic(align_corners=True) 104 -> 416 -> inverse affine warp back to the
original photo (cv2.warpAffine INTER_CUBIC semantics) -> threshold 0.35 -> IoU against the ground-truth mask.

The reference moves every prediction to the host and calls OpenCV per sample; `evaluate_batch` does the whole batch in
two kernel launches (csrc/postproc.cu) and returns the per-sample IoU (and, on request, the binary masks).  Only the
ground-truth masks (uint8, packed) go up and 16 bytes per sample come back.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


class _Sample(C.Structure):
    _fields_ = [("m", C.c_double * 6), ("h", C.c_int32), ("w", C.c_int32), ("off", C.c_int64)]


def evaluate_batch(pred_logits: torch.Tensor, mats: Sequence, ori_sizes: Sequence, gt_masks: Optional[List] = None,
                   input_size: int = 416, threshold: float = 0.35, return_masks: bool = False):
    """pred_logits: [B,1,h,w] or [B,h,w] CUDA float (the model's eval output); mats[b]: the 2x3 matrix the reference passes
    to cv2.warpAffine (`param['inverse']`); ori_sizes[b] = (h, w) of the original photo; gt_masks[b]: uint8 / bool
    [h, w] (non-zero = object; the reference divides the PNG by 255 and uses it as a boolean).
    -> ious float64 [B] (numpy), and the list of uint8 [h, w] prediction masks when return_masks."""
    if not pred_logits.is_cuda:
        raise RuntimeError("cris.pytorch_b200.postproc runs on the GPU only")
    L = _lib.lib()
    if L.cris_postproc_sample_bytes() != C.sizeof(_Sample):
        raise RuntimeError("postproc sample record layout mismatch")
    x = pred_logits.detach().float()
    if x.dim() == 4:
        x = x[:, 0]
    x = x.contiguous()
    B, H, W = x.shape
    dev = x.device
    sizes = [(int(s[0]), int(s[1])) for s in ori_sizes]
    offs, total = [], 0
    for h, w in sizes:
        offs.append(total)
        total += h * w
    recs = (_Sample * B)()
    for b in range(B):
        m = np.asarray(mats[b], dtype=np.float64).reshape(6)
        for i in range(6):
            recs[b].m[i] = float(m[i])
        recs[b].h, recs[b].w, recs[b].off = sizes[b][0], sizes[b][1], offs[b]
    rec_host = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8)
    gt_dev = None
    if gt_masks is not None:
        packed = np.empty(total, np.uint8)
        for b in range(B):
            g = gt_masks[b]
            g = g.cpu().numpy() if isinstance(g, torch.Tensor) else np.asarray(g)
            if g.shape != sizes[b]:
                raise ValueError(f"ground-truth mask {b} has shape {g.shape}, expected {sizes[b]}")
            packed[offs[b]:offs[b] + g.size] = (g.reshape(-1) != 0)
        gt_dev = torch.from_numpy(packed).pin_memory().to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        rec_dev = rec_host.pin_memory().to(dev, non_blocking=True)
        up = torch.empty(B, input_size, input_size, dtype=torch.float32, device=dev)
        counts = torch.zeros(B, 2, dtype=torch.int64, device=dev)
        out = torch.empty(total, dtype=torch.uint8, device=dev) if return_masks else None
        _lib.call("cris_postproc_upsample", x.data_ptr(), up.data_ptr(), B, H, W, input_size, input_size)
        _lib.call("cris_postproc_warp_iou", up.data_ptr(), B, input_size, input_size, rec_dev.data_ptr(),
                  gt_dev.data_ptr() if gt_dev is not None else None, out.data_ptr() if out is not None else None,
                  float(threshold), counts.data_ptr(), max(h * w for h, w in sizes))
        c = counts.cpu().numpy().astype(np.float64)
    ious = c[:, 0] / (c[:, 1] + 1e-6)   # engine/engine.py:120-123
    if return_masks:
        o = out.cpu().numpy()
        return ious, [o[offs[b]:offs[b] + sizes[b][0] * sizes[b][1]].reshape(sizes[b]) for b in range(B)]
    return ious
