"""CPU restatement of the reference's input transform (TEST INFRASTRUCTURE — nothing in cris/ imports it).

The reference's dataset turns a decoded photo into the model's input (utils/dataset.py:136-163,193-221):

    img  = cv2.warpAffine(img_rgb_u8, mat, (416, 416), flags=cv2.INTER_CUBIC, borderValue=[mean * 255])   # letterbox
    mask = cv2.warpAffine(mask_u8,   mat, (416, 416), flags=cv2.INTER_LINEAR, borderValue=0.) / 255.      # train only
    img  = torch.from_numpy(img.transpose(2, 0, 1)).float().div_(255.).sub_(mean).div_(std)              # convert()

The arithmetic lives in OpenCV (imgwarp.cpp: warpAffine + remapBicubic / remapBilinear for 8-bit images), which is not
under /root/reference; cv2 4.13 is installed in the build container, so this restatement is PINNED against the real
thing: tests/test_feeder_cpu.py runs cv2.warpAffine next to these functions on seeded images (bit-exact, both
interpolations) and tests/golden/feeder_r02.npz stores cv2 + torch outputs for the GPU box.

8-bit warpAffine: destination coordinates are mapped in fixed point exactly as for float images (AB_BITS = 10, rounded
to 1/32 pixel), but the interpolation runs in INTEGERS: the 2-D weights are the float outer product of the 1-D kernels
scaled by 2^15 and rounded to short, then ONE central tap is corrected so that the weights sum to exactly 2^15
(initInterTab2D; the search for that tap walks k1, k2 in {ksize/2, ksize/2 + 1} — past the centre, as OpenCV does);
the result is (sum + 2^14) >> 15 saturated to uchar.  With BORDER_CONSTANT a destination pixel whose whole footprint
lies outside the source takes the border value, otherwise outside taps contribute the border value.
"""
from __future__ import annotations

import numpy as np

from .postproc_oracle import AB_BITS, AB_SCALE, INTER_BITS, INTER_TAB_SIZE, _cv_round, cubic_tab, invert_affine

COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS
MEAN = np.array([0.48145466, 0.4578275, 0.40821073], np.float32)   # utils/dataset.py:106-109
STD = np.array([0.26862954, 0.26130258, 0.27577711], np.float32)
BORDER = [0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255]       # utils/dataset.py:154 (doubles)


def linear_tab() -> np.ndarray:
    """initInterTab1D(INTER_LINEAR): 32 x 2 float weights (1 - x, x)."""
    x = (np.arange(INTER_TAB_SIZE, dtype=np.float32) * np.float32(1.0 / INTER_TAB_SIZE)).astype(np.float32)
    return np.stack([np.float32(1) - x, x], 1).astype(np.float32)


def fixed_tab(method: str) -> np.ndarray:
    """initInterTab2D(method, fixpt=true): [32*32, k*k] integer weights, index (fy * 32 + fx), tap (row * k + col)."""
    t = cubic_tab() if method == "cubic" else linear_tab()
    k = t.shape[1]
    out = np.zeros((INTER_TAB_SIZE * INTER_TAB_SIZE, k * k), np.int64)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (t[i][:, None] * t[j][None, :]).astype(np.float32)
            it = np.clip(np.rint(v * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int64).reshape(-1)
            diff = int(it.sum()) - COEF_SCALE
            if diff != 0:
                k2 = k // 2
                big = small = k2 * k + k2

                def at(idx):   # OpenCV reads past the 4-tap (linear) entry here; those reads never decide for cubic
                    return it[idx] if idx < it.size else 0

                for k1 in range(k2, k2 + 2):
                    for kk in range(k2, k2 + 2):
                        idx = k1 * k + kk
                        if at(idx) < at(small):
                            small = idx
                        elif at(idx) > at(big):
                            big = idx
                if diff < 0:
                    it[big] -= diff
                else:
                    it[small] -= diff
            out[i * INTER_TAB_SIZE + j] = it
    return out


_TABS = {}


def _tab(method):
    if method not in _TABS:
        _TABS[method] = fixed_tab(method)
    return _TABS[method]


def warp_affine_u8(src: np.ndarray, M: np.ndarray, w: int, h: int, method: str, border) -> np.ndarray:
    """cv2.warpAffine(src_u8, M, (w, h), flags=INTER_CUBIC | INTER_LINEAR, borderValue=border) for [H, W] or [H, W, C]."""
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    sh, sw, cn = src.shape
    k = 4 if method == "cubic" else 2
    off = 1 if method == "cubic" else 0
    Mi = invert_affine(M)
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    adelta = _cv_round(Mi[0, 0] * xs * AB_SCALE)
    bdelta = _cv_round(Mi[1, 0] * xs * AB_SCALE)
    round_delta = AB_SCALE // INTER_TAB_SIZE // 2
    X0 = _cv_round((Mi[0, 1] * ys + Mi[0, 2]) * AB_SCALE) + round_delta
    Y0 = _cv_round((Mi[1, 1] * ys + Mi[1, 2]) * AB_SCALE) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767) - off
    sy = np.clip(Y >> INTER_BITS, -32768, 32767) - off
    wts = _tab(method)[(Y & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (X & (INTER_TAB_SIZE - 1))]
    border = np.atleast_1d(np.asarray(border, np.float64))
    cval = np.clip(np.rint(border), 0, 255).astype(np.int64)      # saturate_cast<uchar>(double)
    cval = np.resize(cval, cn) if cval.size < cn else cval[:cn]
    outside = (sx >= sw) | (sx + k <= 0) | (sy >= sh) | (sy + k <= 0)
    out = np.zeros((h, w, cn), np.uint8)
    for c in range(cn):
        acc = np.full((h, w), cval[c] * COEF_SCALE, np.int64)
        for i in range(k):
            yi = sy + i
            oky = (yi >= 0) & (yi < sh)
            for j in range(k):
                xj = sx + j
                ok = oky & (xj >= 0) & (xj < sw)
                v = src[np.clip(yi, 0, sh - 1), np.clip(xj, 0, sw - 1), c].astype(np.int64)
                acc += np.where(ok, (v - cval[c]) * wts[:, :, i * k + j], 0)
        val = np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255)
        out[:, :, c] = np.where(outside, cval[c], val)
    return out[:, :, 0] if squeeze else out


def normalise(img_u8: np.ndarray) -> np.ndarray:
    """RefDataset.convert (utils/dataset.py:210-215): HWC uint8 -> CHW float32, ((x / 255) - mean) / std in float32."""
    x = img_u8.transpose(2, 0, 1).astype(np.float32)
    x = (x / np.float32(255.0)).astype(np.float32)
    x = (x - MEAN.reshape(3, 1, 1)).astype(np.float32)
    return (x / STD.reshape(3, 1, 1)).astype(np.float32)


def letterbox(img_rgb_u8: np.ndarray, mat: np.ndarray, mask_u8=None, size: int = 416):
    """One sample of utils/dataset.py:148-163: -> (img float32 [3, size, size], mask float32 [size, size] or None)."""
    warped = warp_affine_u8(img_rgb_u8, mat, size, size, "cubic", BORDER)
    img = normalise(warped)
    mask = None
    if mask_u8 is not None:
        m = warp_affine_u8(mask_u8, mat, size, size, "linear", [0.0])
        mask = (m.astype(np.float64) / 255.0).astype(np.float32)   # numpy uint8 / 255. is float64, then .float()
    return img, mask
