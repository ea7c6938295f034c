"""GPU input transform (the step right before the model in the reference's dataset, utils/dataset.py:148-163,210-221):
letterbox warp of the decoded 8-bit photo (cv2.warpAffine INTER_CUBIC, mean-colour border), the bilinear warp of the
training mask, ToTensor + CLIP normalisation — for a whole batch in ONE kernel launch (csrc/feeder.cu), bit-exact with
cv2 + torch.

The reference runs this per sample on loader workers and ships fp32 tensors to the GPU (2.8 MB per 416x416 sample with
its mask); here the decoded uint8 photos travel (0.9 MB for 480x640, 1.2 MB with the mask) and the model's input is
produced where it is consumed.  JPEG decoding, the LMDB read and the BPE tokenizer stay on the host (not built, see
DESIGN.md section 9).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

MEAN = (0.48145466, 0.4578275, 0.40821073)      # utils/dataset.py:106-109
STD = (0.26862954, 0.26130258, 0.27577711)
BORDER = (0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255)   # utils/dataset.py:154


class _Sample(C.Structure):
    _fields_ = [("m", C.c_double * 6), ("h", C.c_int32), ("w", C.c_int32), ("img_off", C.c_int64), ("mask_off", C.c_int64)]


def letterbox_batch(images: Sequence[np.ndarray], mats: Sequence, masks: Optional[Sequence] = None, input_size: int = 416,
                    device=None, border=BORDER, mean=MEAN, std=STD):
    """images[b]: uint8 [h, w, 3] RGB (what `cv2.cvtColor(cv2.imdecode(...), COLOR_BGR2RGB)` returns); mats[b]: the 2x3
    matrix `getTransformMat` hands to cv2.warpAffine; masks[b]: uint8 [h, w] (0 / 255) or None.
    -> (img float32 [B, 3, S, S] on the device, mask float32 [B, S, S] or None) — the tensors RefDataset.__getitem__
    returns, stacked."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("cris.pytorch_b200.feeder runs on the GPU only")
    L = _lib.lib()
    if L.cris_feeder_sample_bytes() != C.sizeof(_Sample):
        raise RuntimeError("feeder sample record layout mismatch")
    B, S = len(images), int(input_size)
    if B == 0 or len(mats) != B or (masks is not None and len(masks) != B):
        raise ValueError("letterbox_batch: images, mats (and masks) must have the same non-zero length")
    recs = (_Sample * B)()
    img_total = mask_total = 0
    for b, im in enumerate(images):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise ValueError(f"image {b}: expected uint8 [h, w, 3], got {im.dtype} {im.shape}")
        h, w = im.shape[:2]
        m = np.asarray(mats[b], dtype=np.float64).reshape(6)
        for i in range(6):
            recs[b].m[i] = float(m[i])
        recs[b].h, recs[b].w, recs[b].img_off, recs[b].mask_off = h, w, img_total, -1
        img_total += h * w * 3
        if masks is not None and masks[b] is not None:
            mk = masks[b]
            if mk.dtype != np.uint8 or mk.shape != (h, w):
                raise ValueError(f"mask {b}: expected uint8 {(h, w)}, got {mk.dtype} {mk.shape}")
            recs[b].mask_off = mask_total
            mask_total += h * w
    host_img = torch.empty(img_total, dtype=torch.uint8).pin_memory()
    hv = host_img.numpy()
    for b, im in enumerate(images):
        hv[recs[b].img_off:recs[b].img_off + im.size] = im.reshape(-1)
    host_mask = None
    if masks is not None:
        host_mask = torch.empty(max(mask_total, 1), dtype=torch.uint8).pin_memory()
        mv = host_mask.numpy()
        for b, mk in enumerate(masks):
            if mk is not None:
                mv[recs[b].mask_off:recs[b].mask_off + mk.size] = mk.reshape(-1)
    rec_host = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).pin_memory()
    with torch.cuda.device(device):
        d_img = host_img.to(device, non_blocking=True)
        d_mask = host_mask.to(device, non_blocking=True) if host_mask is not None else None
        d_rec = rec_host.to(device, non_blocking=True)
        out = torch.empty(B, 3, S, S, dtype=torch.float32, device=device)
        out_mask = torch.empty(B, S, S, dtype=torch.float32, device=device) if masks is not None else None
        bord = (C.c_double * 3)(*[float(v) for v in border])
        mean_c = (C.c_float * 3)(*[float(v) for v in mean])
        std_c = (C.c_float * 3)(*[float(v) for v in std])
        _lib.call("cris_feeder_letterbox", d_img.data_ptr(), d_mask.data_ptr() if d_mask is not None else None,
                  d_rec.data_ptr(), B, S, S, C.addressof(bord), C.addressof(mean_c), C.addressof(std_c), out.data_ptr(),
                  out_mask.data_ptr() if out_mask is not None else None)
        # the pinned staging buffers must outlive the asynchronous copies
        torch.cuda.current_stream().synchronize()
    return out, out_mask
