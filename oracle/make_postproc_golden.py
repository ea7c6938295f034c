"""Golden vectors for the evaluation post-processing, produced by the REFERENCE'S OWN LINES (torch F.interpolate + cv2
warpAffine, engine/engine.py:101-124) in the build container (torch + cv2 present; the GPU box has no /root/reference
but does not need it for this).   python oracle/make_postproc_golden.py  ->  tests/golden/postproc_r02.npz"""
import os
import sys

import cv2
import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = [(480, 640), (640, 427), (375, 500), (333, 500), (612, 612), (120, 900), (416, 416), (351, 640)]


def transform_mats(ori_h, ori_w, inp=416):
    """utils/dataset.py:193-209 getTransformMat(img_size, inverse=True)"""
    scale = min(inp / ori_h, inp / ori_w)
    new_h, new_w = ori_h * scale, ori_w * scale
    bias_x, bias_y = (inp - new_w) / 2., (inp - new_h) / 2.
    src = np.array([[0, 0], [ori_w, 0], [0, ori_h]], np.float32)
    dst = np.array([[bias_x, bias_y], [new_w + bias_x, bias_y], [bias_x, new_h + bias_y]], np.float32)
    return cv2.getAffineTransform(src, dst), cv2.getAffineTransform(dst, src)


def synth_case(b, oh, ow):
    g = np.random.default_rng(100 + b)
    yy, xx = np.mgrid[0:104, 0:104].astype(np.float32)
    cy, cx = g.uniform(30, 74, 2)
    ry, rx = g.uniform(10, 30, 2)
    blob = 3.0 - 4.0 * (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2)
    logits = (blob + g.standard_normal((104, 104)) * 0.6).astype(np.float32)
    Y, X = np.mgrid[0:oh, 0:ow].astype(np.float32)
    gt = ((((Y - oh * g.uniform(0.35, 0.65)) / (oh * 0.22)) ** 2 + ((X - ow * g.uniform(0.35, 0.65)) / (ow * 0.2)) ** 2) < 1).astype(np.uint8) * 255
    return logits, gt


def main():
    out = {"sizes": np.array(SIZES, np.int32)}
    for b, (oh, ow) in enumerate(SIZES):
        logits, gt = synth_case(b, oh, ow)
        _, mat_inv = transform_mats(oh, ow)
        preds = torch.sigmoid(torch.from_numpy(logits))[None, None]                                   # engine.py:103
        preds = F.interpolate(preds, size=(416, 416), mode="bicubic", align_corners=True).squeeze()   # :104-107
        pred = cv2.warpAffine(preds.numpy(), mat_inv, (ow, oh), flags=cv2.INTER_CUBIC, borderValue=0.)  # :115-117
        pred = np.array(pred > 0.35)                                                                   # :118
        mask = gt / 255.                                                                               # :119-120
        inter = np.logical_and(pred, mask)
        union = np.logical_or(pred, mask)
        iou = np.sum(inter) / (np.sum(union) + 1e-6)                                                   # :121-123
        out[f"logits{b}"] = logits
        out[f"mat{b}"] = mat_inv.astype(np.float64)
        out[f"gt{b}"] = np.packbits(gt != 0)
        out[f"pred{b}"] = np.packbits(pred)
        out[f"iou{b}"] = np.float64(iou)
        print(f"case {b}: {oh}x{ow} iou {iou:.6f} pred px {int(pred.sum())}")
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "postproc_r02.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
