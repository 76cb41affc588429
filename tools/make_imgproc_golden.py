"""Golden vectors for the image kernels, generated from cv2 (the dependency the reference calls) in this container:

    python tools/make_imgproc_golden.py        -> tests/golden/imgproc_cv2.npz

Float resizes are generated with cv2.ipp.setUseIPP(False): the opencv-python wheel otherwise routes float32
INTER_LINEAR through Intel IPP, whose result differs from OpenCV's own algorithm by up to ~1e-4 (both are recorded).
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cv2_reference as R  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    out = {}
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    out["u8_src"] = img
    out["u8_to_32x32"] = cv2.resize(img, (32, 32), interpolation=cv2.INTER_LINEAR)
    out["u8_to_64x48"] = cv2.resize(img, (64, 48), interpolation=cv2.INTER_LINEAR)
    img2 = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    out["u8_src_2x"] = img2
    out["u8_2x_to_32x32"] = cv2.resize(img2, (32, 32), interpolation=cv2.INTER_LINEAR)      # INTER_AREA shortcut
    m = rng.normal(0.3, 1.0, (32, 32)).astype(np.float32)
    out["f32_src"] = m
    out["f32_to_75x41_ipp"] = cv2.resize(m, (75, 41))
    cv2.ipp.setUseIPP(False)
    out["f32_to_75x41"] = cv2.resize(m, (75, 41))
    out["f32_to_16x16"] = cv2.resize(m, (16, 16))
    cv2.ipp.setUseIPP(True)
    mask = R.blob_mask(96, 128, seed=3)
    out["mask"] = mask
    for mode in ("watermark", "text", "mixed"):
        out[f"opt_{mode}"] = R.optimize_mask(mask.copy(), mode)
    n, labels, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=8)
    out["cc_labels"] = labels.astype(np.int32)
    out["cc_stats"] = stats.astype(np.int32)
    out["text_score"] = np.float64(R.analyze_text_features(mask))
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "imgproc_cv2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
