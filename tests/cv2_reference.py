"""The reference's own OpenCV call sequences (reference src/predict.py), executed with the cv2 installed here.

The reference module cannot be imported (top-level ``from iopaint...`` / smp imports), so the bodies of
the functions on the path are restated call for call - same cv2 functions, same arguments, same order.  These are the
PIN for oracle/imgproc_oracle.py and for the CUDA kernels: cv2 is the dependency the reference really calls.
Test infrastructure only.
"""
import cv2
import numpy as np


def optimize_text_mask(mask):                                   # reference src/predict.py:192-230
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2))
    mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, k, iterations=1)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, k, iterations=2)
    mask_h = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 1)), iterations=1)
    mask_v = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (1, 5)), iterations=1)
    mask = cv2.bitwise_or(mask_h, mask_v)
    mask = cv2.dilate(mask, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (4, 4)), iterations=1)
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=8)
    if num_labels > 1:
        out = np.zeros_like(labels, dtype=np.uint8)
        for i in range(1, num_labels):
            if stats[i, cv2.CC_STAT_AREA] > 50:
                out[labels == i] = 255
        mask = out
    return mask


def optimize_watermark_mask(mask):                              # reference src/predict.py:232-273
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, k, iterations=1)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7))
    mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, k, iterations=3)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (11, 11))
    mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, k, iterations=2)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (9, 9))
    mask = cv2.dilate(mask, k, iterations=2)
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=8)
    if num_labels > 1:
        largest = 1 + np.argmax(stats[1:, cv2.CC_STAT_AREA])
        mask = (labels == largest).astype(np.uint8) * 255
        if stats[largest, cv2.CC_STAT_AREA] < 500:
            mask = np.zeros_like(labels, dtype=np.uint8)
            for i in range(1, num_labels):
                if stats[i, cv2.CC_STAT_AREA] > 200:
                    mask[labels == i] = 255
    mask = cv2.GaussianBlur(mask, (3, 3), 0.5)
    _, mask = cv2.threshold(mask, 127, 255, cv2.THRESH_BINARY)
    return mask


def optimize_mixed_mask(mask):                                  # reference src/predict.py:275-301
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2))
    mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, k, iterations=1)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
    mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, k, iterations=2)
    mask = cv2.dilate(mask, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (6, 6)), iterations=1)
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(mask, connectivity=8)
    if num_labels > 1:
        out = np.zeros_like(labels, dtype=np.uint8)
        for i in range(1, num_labels):
            if stats[i, cv2.CC_STAT_AREA] > 100:
                out[labels == i] = 255
        mask = out
    return mask


def optimize_mask(mask, mask_type="watermark"):                 # reference src/predict.py:161-190
    if mask is None:
        return mask
    if len(mask.shape) == 3:
        mask = cv2.cvtColor(mask, cv2.COLOR_BGR2GRAY)
    _, mask = cv2.threshold(mask, 127, 255, cv2.THRESH_BINARY)
    if mask_type == "text":
        return optimize_text_mask(mask)
    if mask_type == "mixed":
        return optimize_mixed_mask(mask)
    return optimize_watermark_mask(mask)


def analyze_text_features(mask_binary):                         # reference src/predict.py:443-508
    if mask_binary is None or np.sum(mask_binary) == 0:
        return 0.0
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(mask_binary, connectivity=8)
    if num_labels <= 1:
        return 0.0
    text_indicators = 0
    total_components = num_labels - 1
    for i in range(1, num_labels):
        area = stats[i, cv2.CC_STAT_AREA]
        width = stats[i, cv2.CC_STAT_WIDTH]
        height = stats[i, cv2.CC_STAT_HEIGHT]
        if area == 0 or width == 0 or height == 0:
            continue
        aspect_ratio = max(width, height) / min(width, height)
        density = area / (width * height)
        score = 0
        if 1 <= aspect_ratio <= 5:
            score += 0.3
        elif 5 < aspect_ratio <= 10:
            score += 0.1
        if 0.3 <= density <= 0.8:
            score += 0.3
        elif 0.2 <= density < 0.3 or 0.8 < density <= 0.9:
            score += 0.1
        if 50 <= area <= 5000:
            score += 0.4
        elif 20 <= area < 50 or 5000 < area <= 10000:
            score += 0.2
        if score > 0.5:
            text_indicators += 1
    if total_components == 0:
        return 0.0
    text_ratio = text_indicators / total_components
    if total_components >= 3 and text_ratio > 0.5:
        return min(text_ratio + 0.2, 1.0)
    return text_ratio


def analyze_ocr_features(image_rgb, mask_binary):               # reference src/predict.py:510-558
    try:
        mask3 = cv2.cvtColor(mask_binary, cv2.COLOR_GRAY2RGB)
        masked = cv2.bitwise_and(image_rgb, mask3)
        gray = cv2.cvtColor(masked, cv2.COLOR_RGB2GRAY)
        edges = cv2.Canny(gray, 50, 150)
        nz = np.sum(mask_binary > 0)
        edge_density = np.sum(edges > 0) / nz if nz > 0 else 0
        gx = cv2.Sobel(gray, cv2.CV_64F, 1, 0, ksize=3)
        gy = cv2.Sobel(gray, cv2.CV_64F, 0, 1, ksize=3)
        angles = np.arctan2(gy, gx)
        region = mask_binary > 0
        angle_variance = np.var(angles[region]) if np.sum(region) > 0 else 0
        s = 0
        if 0.1 <= edge_density <= 0.4:
            s += 0.5
        elif 0.05 <= edge_density < 0.1 or 0.4 < edge_density <= 0.6:
            s += 0.2
        if 1.0 <= angle_variance <= 3.0:
            s += 0.5
        elif 0.5 <= angle_variance < 1.0 or 3.0 < angle_variance <= 4.0:
            s += 0.2
        return min(s, 1.0)
    except Exception:  # noqa: BLE001 - reference behaviour
        return 0.0


def detect_watermark_type(image_rgb, mask_binary):              # reference src/predict.py:414-441
    try:
        total = analyze_text_features(mask_binary) * 0.6 + analyze_ocr_features(image_rgb, mask_binary) * 0.4
        if total > 0.7:
            return "text"
        if total > 0.3:
            return "mixed"
        return "watermark"
    except Exception:  # noqa: BLE001
        return "watermark"


def blob_mask(h, w, seed, n_blobs=6, noise=0.01):
    """A seeded binary test mask: a few ellipses / rectangles / thin strokes plus salt noise (uint8 {0,255})."""
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.uint8)
    for _ in range(n_blobs):
        cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
        ry, rx = int(rng.integers(1, max(2, h // 4))), int(rng.integers(1, max(2, w // 4)))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            cv2.ellipse(m, (cx, cy), (rx, ry), float(rng.integers(0, 180)), 0, 360, 255, -1)
        elif kind == 1:
            cv2.rectangle(m, (cx, cy), (min(w - 1, cx + rx), min(h - 1, cy + ry)), 255, -1)
        else:
            cv2.line(m, (cx, cy), (int(rng.integers(0, w)), int(rng.integers(0, h))), 255, int(rng.integers(1, 4)))
    m[rng.random((h, w)) < noise] = 255
    return m
