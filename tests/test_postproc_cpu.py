"""CPU: the restatement of the reference's evaluation post-processing (oracle/postproc_oracle.py) against
(a) the committed vectors produced by the reference's own lines — torch F.interpolate + cv2.warpAffine,
    engine/engine.py:101-124 — (tests/golden/postproc_r02.npz, oracle/make_postproc_golden.py), and
(b) when cv2 is importable, the live torch + cv2 pipeline on fresh seeded inputs (odd sizes, near-degenerate matrices)."""
import os

import numpy as np
import pytest

from oracle import postproc_oracle as P


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, "postproc_r02.npz"))


def test_oracle_reproduces_reference_postprocessing(golden_dir):
    g = _load(golden_dir)
    for b, (oh, ow) in enumerate(g["sizes"]):
        gt = np.unpackbits(g[f"gt{b}"])[:oh * ow].reshape(oh, ow)
        ref = np.unpackbits(g[f"pred{b}"])[:oh * ow].reshape(oh, ow).astype(bool)
        pred, iou = P.postprocess(g[f"logits{b}"], g[f"mat{b}"], (oh, ow), gt)
        flips = int((pred != ref).sum())
        assert flips <= max(2, int(2e-5 * oh * ow)), (b, flips)          # float summation order at the 0.35 boundary
        assert abs(iou - float(g[f"iou{b}"])) <= 1e-4


def test_oracle_matches_live_torch_and_cv2():
    cv2 = pytest.importorskip("cv2")
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(7)
    for oh, ow in [(97, 211), (500, 333), (416, 416), (33, 700)]:
        logits = (rng.standard_normal((104, 104)) * 2).astype(np.float32)
        up_ref = F.interpolate(torch.sigmoid(torch.from_numpy(logits))[None, None], size=(416, 416), mode="bicubic",
                               align_corners=True).squeeze().numpy()
        prob = (1 / (1 + np.exp(-logits))).astype(np.float32)
        assert np.abs(P.bicubic_upsample_align_corners(prob, 416, 416) - up_ref).max() <= 4e-6
        scale = min(416 / oh, 416 / ow)
        bx, by = (416 - ow * scale) / 2., (416 - oh * scale) / 2.
        src = np.array([[0, 0], [ow, 0], [0, oh]], np.float32)
        dst = np.array([[bx, by], [ow * scale + bx, by], [bx, oh * scale + by]], np.float32)
        mat_inv = cv2.getAffineTransform(dst, src)
        # a slightly rotated / sheared matrix too: the fixed-point path must hold for general affine maps
        for M in (mat_inv, mat_inv @ np.array([[0.98, 0.05, 3.0], [-0.04, 1.01, -2.0], [0, 0, 1.0]])):
            wr = cv2.warpAffine(up_ref, M, (ow, oh), flags=cv2.INTER_CUBIC, borderValue=0.)
            wo = P.warp_affine_cubic(up_ref, M, ow, oh)
            assert np.abs(wo - wr).max() <= 2e-6
            assert int(((wo > 0.35) != (wr > 0.35)).sum()) <= 2
