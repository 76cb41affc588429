"""oracle/imgproc_oracle.py pinned against cv2 - the dependency the reference's pre/post-processing really calls
(reference src/predict.py:588-664, :161-301, :443-508) - and against the committed cv2 golden vectors."""
import os

import cv2
import numpy as np
import pytest

from oracle import imgproc_oracle as I
from tests import cv2_reference as R

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imgproc_cv2.npz"))


def test_structuring_elements_equal_cv2():
    for shape in (cv2.MORPH_RECT, cv2.MORPH_CROSS, cv2.MORPH_ELLIPSE):
        for kx in range(1, 14):
            for ky in range(1, 14):
                assert np.array_equal(I.structuring_element(shape, (kx, ky)), cv2.getStructuringElement(shape, (kx, ky)))


def test_resize_u8_bit_exact_with_cv2():
    rng = np.random.default_rng(0)
    cases = [(37, 53, 32, 32, 3), (64, 64, 32, 32, 3), (100, 60, 50, 30, 1), (64, 40, 32, 32, 3), (5, 7, 64, 64, 3),
             (1, 1, 16, 16, 3), (2, 3, 1, 1, 1), (333, 517, 128, 128, 3), (128, 128, 128, 128, 3), (90, 31, 256, 200, 3)]
    for _ in range(40):
        sh, sw, dh, dw = (int(v) for v in rng.integers(1, 300, 4))
        cases.append((sh, sw, dw, dh, int(rng.choice([1, 3]))))
    for sh, sw, dw, dh, cn in cases:
        img = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
        img = img[:, :, 0] if cn == 1 else img
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(I.resize_linear_u8(img, dw, dh), want), (sh, sw, dw, dh, cn)


def test_resize_f32_bit_exact_with_opencv_algorithm_and_close_to_ipp():
    rng = np.random.default_rng(1)
    cases = [(32, 32, 75, 41), (32, 32, 16, 16), (64, 64, 640, 480), (64, 64, 33, 97), (16, 16, 16, 16), (64, 64, 32, 32)]
    for sh, sw, dw, dh in cases:
        m = rng.normal(0.5, 1.0, (sh, sw)).astype(np.float32)
        got = I.resize_linear_f32(m, dw, dh)
        ipp = cv2.resize(m, (dw, dh))                    # whatever this cv2 build does (IPP in the pip wheel)
        cv2.ipp.setUseIPP(False)
        try:
            own = cv2.resize(m, (dw, dh))                # OpenCV's own algorithm
        finally:
            cv2.ipp.setUseIPP(True)
        if I.is_area_2x(sw, sh, dw, dh):     # INTER_AREA shortcut: OpenCV's scalar tail columns add in another order (1 ulp)
            assert np.abs(got - own).max() <= 2.4e-7 * np.abs(m).max() and (got != own).mean() < 0.5
        else:
            assert np.array_equal(got, own), (sh, sw, dw, dh)
        assert np.abs(got - ipp).max() <= 5e-4 * max(1.0, np.abs(m).max())
        assert ((got > 0.5) != (ipp > 0.5)).mean() <= 1e-3


def test_morphology_equals_cv2():
    rng = np.random.default_rng(2)
    for it in range(25):
        h, w = (int(v) for v in rng.integers(3, 70, 2))
        m = (rng.random((h, w)) < rng.choice([0.1, 0.5, 0.85])).astype(np.uint8) * 255
        shape = int(rng.choice([cv2.MORPH_RECT, cv2.MORPH_ELLIPSE, cv2.MORPH_CROSS]))
        ks = [(2, 2), (3, 3), (4, 4), (5, 5), (6, 6), (7, 7), (9, 9), (11, 11), (5, 1), (1, 5)][it % 10]
        iters = 1 + it % 3
        el, cel = I.structuring_element(shape, ks), cv2.getStructuringElement(shape, ks)
        assert np.array_equal(I.erode(m, el, iters), cv2.erode(m, cel, iterations=iters))
        assert np.array_equal(I.dilate(m, el, iters), cv2.dilate(m, cel, iterations=iters))
        assert np.array_equal(I.morph_open(m, el, iters), cv2.morphologyEx(m, cv2.MORPH_OPEN, cel, iterations=iters))
        assert np.array_equal(I.morph_close(m, el, iters), cv2.morphologyEx(m, cv2.MORPH_CLOSE, cel, iterations=iters))


def test_connected_components_labels_and_order_equal_cv2():
    rng = np.random.default_rng(3)
    masks = [(rng.random((int(h), int(w))) < p).astype(np.uint8) * 255
             for h, w, p in zip(rng.integers(1, 48, 40), rng.integers(1, 48, 40), rng.choice([0.2, 0.4, 0.6], 40))]
    # label order differs from first-pixel raster order here: the component at (1,0) lies in block row 0, so it is
    # numbered before the component whose first pixel (0,3) comes earlier in raster order?  cv2 decides.
    crafted = np.zeros((4, 6), np.uint8)
    crafted[1, 0] = 255; crafted[0, 3] = 255; crafted[3, 5] = 255; crafted[2, 2] = 255
    masks += [crafted, np.zeros((5, 5), np.uint8), np.full((4, 7), 255, np.uint8), R.blob_mask(120, 90, 5, noise=0.05)]
    for m in masks:
        n, labels, stats, _ = cv2.connectedComponentsWithStats(m, connectivity=8)
        la, sa = I.connected_components_8(m)
        assert np.array_equal(la, labels)
        assert np.array_equal(sa[1:], stats[1:])


@pytest.mark.parametrize("mode", ["watermark", "text", "mixed"])
def test_optimize_mask_equals_reference_cv2_sequence(mode):
    for seed, (h, w), noise in ((0, (96, 128), 0.01), (1, (70, 200), 0.03), (2, (33, 47), 0.0), (9, (150, 150), 0.002)):
        m = R.blob_mask(h, w, seed, noise=noise)
        assert np.array_equal(I.optimize_mask(m, mode), R.optimize_mask(m.copy(), mode)), (mode, seed)
    # the 'largest < 500 -> keep > 200' branch and the empty mask
    small = np.zeros((80, 80), np.uint8)
    small[5:8, 5:9] = 255; small[40:44, 40:45] = 255
    assert np.array_equal(I.optimize_mask(small, mode), R.optimize_mask(small.copy(), mode))
    z = np.zeros((20, 30), np.uint8)
    assert np.array_equal(I.optimize_mask(z, mode), R.optimize_mask(z.copy(), mode))
    grey = (np.arange(64 * 64).reshape(64, 64) % 256).astype(np.uint8)        # non-binary input: threshold 127 first
    assert np.array_equal(I.optimize_mask(grey, mode), R.optimize_mask(grey.copy(), mode))


def test_gaussian_threshold_is_identity_on_binary_masks():
    for seed in range(4):
        m = R.blob_mask(64, 80, seed, noise=0.2)
        b = cv2.GaussianBlur(m, (3, 3), 0.5)
        _, b = cv2.threshold(b, 127, 255, cv2.THRESH_BINARY)
        assert np.array_equal(b, m)


def test_text_features_equal_reference():
    for seed in range(6):
        m = R.blob_mask(100, 140, seed, n_blobs=3 + seed, noise=0.01 * seed)
        assert I.analyze_text_features(m) == R.analyze_text_features(m)
    assert I.analyze_text_features(np.zeros((8, 8), np.uint8)) == 0.0


def test_golden_vectors():
    g = GOLD
    assert np.array_equal(I.resize_linear_u8(g["u8_src"], 32, 32), g["u8_to_32x32"])
    assert np.array_equal(I.resize_linear_u8(g["u8_src"], 64, 48), g["u8_to_64x48"])
    assert np.array_equal(I.resize_linear_u8(g["u8_src_2x"], 32, 32), g["u8_2x_to_32x32"])
    assert np.array_equal(I.resize_linear_f32(g["f32_src"], 75, 41), g["f32_to_75x41"])
    assert np.abs(I.resize_linear_f32(g["f32_src"], 16, 16) - g["f32_to_16x16"]).max() <= 2.4e-7 * np.abs(g["f32_src"]).max()
    assert np.abs(I.resize_linear_f32(g["f32_src"], 75, 41) - g["f32_to_75x41_ipp"]).max() < 1e-3
    for mode in ("watermark", "text", "mixed"):
        assert np.array_equal(I.optimize_mask(g["mask"], mode), g[f"opt_{mode}"])
    la, sa = I.connected_components_8(g["mask"])
    assert np.array_equal(la, g["cc_labels"]) and np.array_equal(sa[1:], g["cc_stats"][1:])
    assert I.analyze_text_features(g["mask"]) == float(g["text_score"])
