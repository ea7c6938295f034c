"""CPU restatement of the reference's evaluation post-processing (TEST INFRASTRUCTURE — nothing in cris/ imports it).

The reference's validate()/inference() (engine/engine.py:101-124,172-190) turn the model's logits into an IoU:

    pred = sigmoid(logits)                                                      # [B,1,104,104]
    pred = F.interpolate(pred, size=(416,416), mode='bicubic', align_corners=True)
    pred = cv2.warpAffine(pred[b], mat_inv[b], (w, h), flags=cv2.INTER_CUBIC, borderValue=0.)   # back to the photo
    pred = pred > 0.35 ; iou = sum(pred & gt) / (sum(pred | gt) + 1e-6)

The arithmetic lives in two third-party libraries that are not under /root/reference: PyTorch (ATen
upsample_bicubic2d, UpSampleKernel.cpp) and OpenCV (imgwarp.cpp warpAffine + remapBicubic).  Both are installed in the
build container (torch 2.11, cv2 4.13), so this restatement is PINNED against the real thing:
tests/test_postproc_cpu.py runs the reference's own lines (torch + cv2) next to these functions on seeded inputs
(bicubic: <= 2e-6; warp: identical thresholded masks), and tests/golden/postproc_*.npz stores vectors produced by
torch + cv2 for the GPU box, where the test compares the CUDA kernels against them.

ATen bicubic, align_corners=True: src = dst * (in-1)/(out-1); taps floor(src)-1..+2 clamped to the image, cubic
convolution coefficients with A = -0.75.
OpenCV warpAffine(INTER_CUBIC, BORDER_CONSTANT 0): M is inverted in double precision, destination coordinates are
mapped in FIXED POINT (AB_BITS = 10, rounded to 1/32 pixel: INTER_BITS = 5), the 4x4 weights are the outer product of
two float cubic kernels (A = -0.75) taken at the 1/32 sub-pixel positions, taps outside the source contribute 0.
"""
from __future__ import annotations

import numpy as np

A = np.float32(-0.75)
AB_BITS, INTER_BITS = 10, 5
AB_SCALE, INTER_TAB_SIZE = 1 << AB_BITS, 1 << INTER_BITS


def _cc1(x):
    return ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + np.float32(1)


def _cc2(x):
    return ((A * x - np.float32(5) * A) * x + np.float32(8) * A) * x - np.float32(4) * A


def bicubic_coeffs(t: np.ndarray) -> np.ndarray:
    """ATen get_cubic_upsample_coefficients (UpSample.h), float32; t in [0,1) -> [..., 4]."""
    t = t.astype(np.float32)
    x2 = np.float32(1) - t
    return np.stack([_cc2(t + np.float32(1)), _cc1(t), _cc1(x2), _cc2(x2 + np.float32(1))], -1).astype(np.float32)


def bicubic_upsample_align_corners(src: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """F.interpolate(mode='bicubic', align_corners=True) on one [H, W] float32 image (engine/engine.py:104-107)."""
    src = src.astype(np.float32)
    H, W = src.shape

    def axis(n_in, n_out):
        scale = np.float32((n_in - 1) / (n_out - 1)) if n_out > 1 else np.float32(0)
        real = scale * np.arange(n_out, dtype=np.float32)
        i0 = np.floor(real).astype(np.int64)
        w = bicubic_coeffs(real - i0.astype(np.float32))
        idx = np.clip(i0[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)
        return idx, w

    iy, wy = axis(H, out_h)
    ix, wx = axis(W, out_w)
    # horizontal pass then vertical pass, float32 accumulation in tap order (ATen's separable kernel)
    tmp = np.zeros((H, out_w), np.float32)
    for j in range(4):
        tmp += src[:, ix[:, j]] * wx[None, :, j]
    out = np.zeros((out_h, out_w), np.float32)
    for i in range(4):
        out += tmp[iy[:, i], :] * wy[:, i, None]
    return out


def _cv_round(x: np.ndarray) -> np.ndarray:
    """cv::saturate_cast<int>(double) = cvRound: round half to even."""
    return np.rint(x).astype(np.int64)


def invert_affine(M: np.ndarray) -> np.ndarray:
    """imgwarp.cpp warpAffine, the !WARP_INVERSE_MAP branch (double precision)."""
    M = np.array(M, np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def cubic_tab() -> np.ndarray:
    """initInterTab1D(INTER_CUBIC): 32 x 4 float weights (interpolateCubic, A = -0.75)."""
    x = (np.arange(INTER_TAB_SIZE, dtype=np.float32) * np.float32(1.0 / INTER_TAB_SIZE)).astype(np.float32)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c0, c1, c2 = c0.astype(np.float32), c1.astype(np.float32), c2.astype(np.float32)
    c3 = (np.float32(1) - c0 - c1 - c2).astype(np.float32)
    return np.stack([c0, c1, c2, c3], 1)


def warp_affine_cubic(src: np.ndarray, M: np.ndarray, w: int, h: int) -> np.ndarray:
    """cv2.warpAffine(src, M, (w, h), flags=cv2.INTER_CUBIC, borderValue=0.) for a float32 single-channel image."""
    src = src.astype(np.float32)
    sh, sw = src.shape
    Mi = invert_affine(M)
    xs = np.arange(w, dtype=np.float64)
    adelta = _cv_round(Mi[0, 0] * xs * AB_SCALE)
    bdelta = _cv_round(Mi[1, 0] * xs * AB_SCALE)
    round_delta = AB_SCALE // INTER_TAB_SIZE // 2
    ys = np.arange(h, dtype=np.float64)
    X0 = _cv_round((Mi[0, 1] * ys + Mi[0, 2]) * AB_SCALE) + round_delta
    Y0 = _cv_round((Mi[1, 1] * ys + Mi[1, 2]) * AB_SCALE) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767) - 1   # saturate_cast<short>, then the top-left tap
    sy = np.clip(Y >> INTER_BITS, -32768, 32767) - 1
    fx, fy = X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)
    tab = cubic_tab()
    out = np.zeros((h, w), np.float32)
    pad = np.zeros((sh + 8, sw + 8), np.float32)   # taps outside the source read the border value 0
    pad[4:4 + sh, 4:4 + sw] = src
    inside = (sx + 3 >= 0) & (sx < sw) & (sy + 3 >= 0) & (sy < sh)
    cx = np.clip(sx, -4, sw) + 4
    cy = np.clip(sy, -4, sh) + 4
    for r in range(4):      # remapBicubic: sum over rows of (4-tap dot product), float accumulation in this order
        wyr = tab[fy, r]
        row = np.zeros((h, w), np.float32)
        for c in range(4):
            wgt = (wyr * tab[fx, c]).astype(np.float32)          # initInterTab2D: float product of the 1-D weights
            row = (row + pad[np.clip(cy + r, 0, sh + 7), np.clip(cx + c, 0, sw + 7)] * wgt).astype(np.float32)
        out = (out + row).astype(np.float32)
    return np.where(inside, out, np.float32(0)).astype(np.float32)


def postprocess(logits: np.ndarray, mat_inv: np.ndarray, ori_hw, gt_mask: np.ndarray, size: int = 416, thr: float = 0.35):
    """One sample of engine/engine.py:101-124: logits [104,104] -> (binary prediction [h,w], IoU vs gt_mask [h,w])."""
    prob = (np.float32(1) / (np.float32(1) + np.exp(-logits.astype(np.float32)))).astype(np.float32)
    up = bicubic_upsample_align_corners(prob, size, size)
    h, w = int(ori_hw[0]), int(ori_hw[1])
    warped = warp_affine_cubic(up, mat_inv, w, h)
    pred = warped > np.float32(thr)
    gt = gt_mask.astype(bool) if gt_mask.dtype != bool else gt_mask
    inter = np.logical_and(pred, gt).sum()
    union = np.logical_or(pred, gt).sum()
    return pred, float(inter / (union + 1e-6))
