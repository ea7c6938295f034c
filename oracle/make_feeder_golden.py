"""Writes tests/golden/feeder_r02.npz: inputs and outputs of the reference's input transform (utils/dataset.py:136-163,
193-221) produced by the REAL libraries (cv2.warpAffine + torch), for the GPU box where neither /root/reference nor the
comparison against cv2 is available.  Run in the build container: python -m oracle.make_feeder_golden"""
import os

import cv2
import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "feeder_r02.npz")
MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073]).reshape(3, 1, 1)
STD = torch.tensor([0.26862954, 0.26130258, 0.27577711]).reshape(3, 1, 1)


def transform_mat(img_size, inp):
    """RefDataset.getTransformMat (utils/dataset.py:193-208)."""
    ori_h, ori_w = img_size
    scale = min(inp / ori_h, inp / ori_w)
    new_h, new_w = ori_h * scale, ori_w * scale
    bias_x, bias_y = (inp - new_w) / 2., (inp - new_h) / 2.
    src = np.array([[0, 0], [ori_w, 0], [0, ori_h]], np.float32)
    dst = np.array([[bias_x, bias_y], [new_w + bias_x, bias_y], [bias_x, new_h + bias_y]], np.float32)
    return cv2.getAffineTransform(src, dst)


def photo(rng, h, w):
    """A photo-like image: smooth colour gradients + blobs + noise (compresses, and has edges for the cubic overshoot)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        img[:, :, c] = 128 + 100 * np.sin(xx / rng.uniform(8, 40) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(8, 40))
    for _ in range(6):
        cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(5, 40)
        img[((yy - cy) ** 2 + (xx - cx) ** 2) < r * r] = rng.uniform(0, 255, 3)
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def blob_mask(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cy, cx = rng.uniform(0.3, 0.7) * h, rng.uniform(0.3, 0.7) * w
    return ((((yy - cy) / (0.25 * h)) ** 2 + ((xx - cx) / (0.3 * w)) ** 2) < 1).astype(np.uint8) * 255


def main():
    rng = np.random.default_rng(20260924)
    out = {}
    cases = [(120, 160, 96), (201, 97, 96), (64, 64, 96), (333, 500, 416)]   # (h, w, input size): landscape, portrait, up-scaling, real size
    for i, (h, w, S) in enumerate(cases):
        img, mask = photo(rng, h, w), blob_mask(rng, h, w)
        mat = transform_mat((h, w), S)
        wi = cv2.warpAffine(img, mat, (S, S), flags=cv2.INTER_CUBIC,
                            borderValue=[0.48145466 * 255, 0.4578275 * 255, 0.40821073 * 255])
        wm = cv2.warpAffine(mask, mat, (S, S), flags=cv2.INTER_LINEAR, borderValue=0.)
        t = torch.from_numpy(wi.transpose((2, 0, 1))).float()
        t.div_(255.).sub_(MEAN).div_(STD)
        m = torch.from_numpy(wm / 255.).float()
        out[f"img{i}"], out[f"mask{i}"], out[f"mat{i}"], out[f"size{i}"] = img, mask, mat, np.int64(S)
        out[f"warped{i}"], out[f"wmask{i}"] = wi, wm
        if S <= 96:   # the float tensors of the small cases; the 416 case is checked through `warped` + the same formula
            out[f"tensor{i}"], out[f"tmask{i}"] = t.numpy(), m.numpy()
    out["n"] = np.int64(len(cases))
    out["versions"] = np.array([f"cv2 {cv2.__version__}", f"torch {torch.__version__}"])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
