"""GPU: the batched input-transform kernel (csrc/feeder.cu) against vectors produced by cv2.warpAffine + torch running
the reference's lines (utils/dataset.py:148-163,210-221; tests/golden/feeder_r02.npz) and against the pinned oracle on
seeded inputs.  Integer interpolation + separately rounded fp32 normalisation: the bar is BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "feeder_r02.npz")


def test_letterbox_matches_cv2_and_torch_vectors():
    from cris.pytorch_b200.feeder import letterbox_batch
    from oracle import feeder_oracle as fo
    g = np.load(GOLD)
    for S in (96, 416):
        idx = [i for i in range(int(g["n"])) if int(g[f"size{i}"]) == S]
        imgs = [g[f"img{i}"] for i in idx]
        masks = [g[f"mask{i}"] for i in idx]
        if len(idx) > 1:
            masks[1] = None                      # a sample without a mask inside a batch that has masks
        out, om = letterbox_batch(imgs, [g[f"mat{i}"] for i in idx], masks, input_size=S)
        out, om = out.cpu().numpy(), om.cpu().numpy()
        for k, i in enumerate(idx):
            want = g[f"tensor{i}"] if f"tensor{i}" in g else fo.normalise(g[f"warped{i}"])
            assert np.array_equal(out[k], want), f"case {i}: image tensor differs ({np.abs(out[k] - want).max()})"
            if masks[k] is None:
                assert not om[k].any()
            else:
                wantm = g[f"tmask{i}"] if f"tmask{i}" in g else (g[f"wmask{i}"].astype(np.float64) / 255.0).astype(np.float32)
                assert np.array_equal(om[k], wantm), f"case {i}: mask differs"


def test_letterbox_matches_oracle_on_random_batches():
    from cris.pytorch_b200.feeder import letterbox_batch
    from oracle import feeder_oracle as fo
    rng = np.random.default_rng(11)
    imgs, mats, masks = [], [], []
    S = 128
    for b in range(6):
        h, w = int(rng.integers(20, 260)), int(rng.integers(20, 260))
        imgs.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        masks.append((rng.integers(0, 2, (h, w), dtype=np.uint8) * 255))
        sc = min(S / h, S / w)
        m = np.array([[sc, 0, (S - w * sc) / 2], [0, sc, (S - h * sc) / 2]], np.float64)
        if b >= 4:   # rotation + shear: the general path of the same kernel
            m = m @ np.array([[0.9, -0.3, 4.0], [0.25, 1.1, -3.0], [0, 0, 1]])
        mats.append(m)
    out, om = letterbox_batch(imgs, mats, masks, input_size=S)
    img_only, none = letterbox_batch(imgs[:2], mats[:2], None, input_size=S)
    assert none is None and torch.equal(img_only, out[:2])
    out, om = out.cpu().numpy(), om.cpu().numpy()
    for b in range(6):
        wi, wm = fo.letterbox(imgs[b], mats[b], masks[b], S)
        assert np.array_equal(out[b], wi), b
        assert np.array_equal(om[b], wm), b


def test_letterbox_rejects_bad_input():
    from cris.pytorch_b200.feeder import letterbox_batch
    with pytest.raises(ValueError):
        letterbox_batch([np.zeros((4, 4), np.uint8)], [np.eye(2, 3)])
    with pytest.raises(RuntimeError):
        letterbox_batch([np.zeros((4, 4, 3), np.uint8)], [np.eye(2, 3)], device="cpu")
