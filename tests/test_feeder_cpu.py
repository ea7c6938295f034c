"""CPU: the input-transform oracle (oracle/feeder_oracle.py) is pinned against the real libraries — cv2.warpAffine on
8-bit images (bit-exact, INTER_CUBIC 3-channel with the mean border and INTER_LINEAR 1-channel) when cv2 is importable,
and always against tests/golden/feeder_r02.npz, which oracle/make_feeder_golden.py wrote from cv2 + torch running the
reference's own lines (utils/dataset.py:148-163,210-221)."""
import os

import numpy as np
import pytest

from oracle import feeder_oracle as fo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "feeder_r02.npz")


def test_fixed_point_tables_sum_to_one():
    for method, k in (("cubic", 16), ("linear", 4)):
        t = fo.fixed_tab(method)
        assert t.shape == (1024, k) and (t.sum(1) == fo.COEF_SCALE).all()
        assert t.min() >= -32768 and t.max() <= 32767
    assert fo.fixed_tab("linear")[0].tolist() == [32767, 0, 0, 1] or fo.fixed_tab("linear")[0].sum() == 32768


def test_oracle_matches_golden_vectors():
    g = np.load(GOLD)
    for i in range(int(g["n"])):
        S = int(g[f"size{i}"])
        img, mask = fo.letterbox(g[f"img{i}"], g[f"mat{i}"], g[f"mask{i}"], S)
        warped = fo.warp_affine_u8(g[f"img{i}"], g[f"mat{i}"], S, S, "cubic", fo.BORDER)
        assert np.array_equal(warped, g[f"warped{i}"]), f"case {i}: cubic warp differs from cv2"
        assert np.array_equal(fo.warp_affine_u8(g[f"mask{i}"], g[f"mat{i}"], S, S, "linear", [0.0]), g[f"wmask{i}"])
        if f"tensor{i}" in g:
            assert np.array_equal(img, g[f"tensor{i}"]), f"case {i}: normalised tensor differs from torch"
            assert np.array_equal(mask, g[f"tmask{i}"])


def test_oracle_matches_cv2_on_random_images():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for trial in range(8):
        h, w = int(rng.integers(30, 180)), int(rng.integers(30, 180))
        S = int(rng.choice([64, 96, 130]))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mask = (rng.integers(0, 2, (h, w), dtype=np.uint8) * 255)
        if trial % 2:
            img = cv2.GaussianBlur(img, (5, 5), 0)
        sc = min(S / h, S / w)
        nh, nw = h * sc, w * sc
        bx, by = (S - nw) / 2, (S - nh) / 2
        mat = cv2.getAffineTransform(np.array([[0, 0], [w, 0], [0, h]], np.float32),
                                     np.array([[bx, by], [nw + bx, by], [bx, nh + by]], np.float32))
        if trial >= 6:   # a general affine map (rotation + shear), not what the dataset produces but the same code path
            mat = mat @ np.array([[0.9, -0.3, 4.0], [0.25, 1.1, -3.0], [0, 0, 1]])
        ref = cv2.warpAffine(img, mat, (S, S), flags=cv2.INTER_CUBIC, borderValue=fo.BORDER)
        assert np.array_equal(fo.warp_affine_u8(img, mat, S, S, "cubic", fo.BORDER), ref), trial
        refm = cv2.warpAffine(mask, mat, (S, S), flags=cv2.INTER_LINEAR, borderValue=0.)
        assert np.array_equal(fo.warp_affine_u8(mask, mat, S, S, "linear", [0.0]), refm), trial
