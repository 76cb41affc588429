"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU (numpy, integer / fixed-point) restatement of the OpenCV operations the reference runs on either side of the
network in ``WatermarkPredictor.step1_batch_predict_watermark_masks`` (reference src/predict.py:588-664):

  * ``cv2.resize(uint8 RGB, (S,S), INTER_LINEAR)``   albumentations ``A.Resize`` of get_val_transform
                                                     (reference src/utils/dataset.py:389-395; predict.py:598-602)
  * ``cv2.resize(float32 mask, (W0,H0))``            reference src/predict.py:620-621
  * ``_optimize_mask`` family                        reference src/predict.py:161-301 (threshold, elliptical
                                                     open / close / dilate, 8-connected components, area filters,
                                                     3x3 Gaussian + threshold)
  * ``_analyze_text_features``                       reference src/predict.py:443-508

PINNED: unlike the network oracle, every function here is checked against the dependency the reference actually
calls - ``cv2`` (opencv-python-headless 4.13, installed in this image and on the GPU box) - bit for bit on seeded
random inputs and on crafted edge cases (tests/test_imgproc_oracle.py), and small golden vectors generated from
cv2 are committed under tests/golden/ (tools/make_imgproc_golden.py).

The restatements follow OpenCV's published algorithms (modules/imgproc/src/resize.cpp: fixed-point linear
resize with INTER_RESIZE_COEF_BITS = 11, the 2x2 INTER_AREA shortcut for exact 2x down-scaling;
morph.dispatch.cpp / filterengine: constant-border erosion / dilation; connectedcomponents.cpp: label order of the
block-based 8-connectivity scan).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


# ------------------------------------------------------------------------------------------------------------
# cv2.resize, INTER_LINEAR
# ------------------------------------------------------------------------------------------------------------
def linear_taps(ssize: int, dsize: int):
    """Source index and fractional weight of every destination coordinate along one axis, as resize.cpp computes
    them: scale = 1/(dsize/ssize) in double, f = float((d + 0.5) * scale - 0.5), s = floor(f), f -= s.
    Returns (s [dsize] int32, f [dsize] float32) BEFORE any clamping."""
    inv_scale = float(dsize) / float(ssize)
    scale = 1.0 / inv_scale
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _x_taps(ssize: int, dsize: int):
    """Horizontal taps with resize.cpp's clamping: s < 0 -> (0, f = 0); s >= ssize - 1 -> (ssize - 1, f = 0)."""
    s, f = linear_taps(ssize, dsize)
    lo = s < 0
    s = np.where(lo, 0, s); f = np.where(lo, np.float32(0), f)
    hi = s >= ssize - 1
    s = np.where(hi, ssize - 1, s); f = np.where(hi, np.float32(0), f)
    return s.astype(np.int32), f.astype(np.float32)


def _round_short(x: np.ndarray) -> np.ndarray:
    """saturate_cast<short>(float): round half to even."""
    return np.clip(np.rint(x.astype(np.float32)), -32768, 32767).astype(np.int32)


def is_area_2x(sw: int, sh: int, dw: int, dh: int) -> bool:
    """cv::resize switches INTER_LINEAR to the fast INTER_AREA path when both scales are exactly 2."""
    return sw == 2 * dw and sh == 2 * dh


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HxW or HxWxC, bit-exact."""
    squeeze = img.ndim == 2
    src = img[:, :, None] if squeeze else img
    sh, sw, cn = src.shape
    if (sw, sh) == (dw, dh):
        return img.copy()
    if is_area_2x(sw, sh, dw, dh):
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        out = out.astype(np.uint8)
        return out[:, :, 0] if squeeze else out
    sx, fx = _x_taps(sw, dw)
    a0 = _round_short((np.float32(1) - fx) * np.float32(COEF_SCALE))
    a1 = _round_short(fx * np.float32(COEF_SCALE))
    sx1 = np.minimum(sx + 1, sw - 1)
    sy, fy = linear_taps(sh, dh)
    b0 = _round_short((np.float32(1) - fy) * np.float32(COEF_SCALE))
    b1 = _round_short(fy * np.float32(COEF_SCALE))
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    s32 = src.astype(np.int32)
    # horizontal pass: int rows scaled by 2^11 (columns at/after xmax read S[sx] * 2048: a1 is 0 there)
    rows = s32[:, sx, :] * a0[None, :, None] + s32[:, sx1, :] * a1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = ((((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2)
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_linear_f32(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(float32 HxW, (dw, dh)) (INTER_LINEAR): float taps, horizontal then vertical pass.  OpenCV's
    SIMD kernels may contract the two multiply-adds, so agreement with cv2 is to the last ulp or two, not bitwise."""
    sh, sw = img.shape
    if (sw, sh) == (dw, dh):
        return img.copy()
    src = img.astype(np.float32)
    if is_area_2x(sw, sh, dw, dh):
        # OpenCV's vector body adds (a + b) + (c + d); its scalar tail (the last few columns, how many depends on the
        # SIMD width of the build) adds ((a + b) + c) + d, so the last ulp of those columns is build-dependent
        return (((src[0::2, 0::2] + src[0::2, 1::2]) + (src[1::2, 0::2] + src[1::2, 1::2])) * np.float32(0.25)).astype(np.float32)
    sx, fx = _x_taps(sw, dw)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy, fy = linear_taps(sh, dh)
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    rows = src[:, sx] * (np.float32(1) - fx)[None, :] + src[:, sx1] * fx[None, :]
    out = rows[y0] * (np.float32(1) - fy)[:, None] + rows[y1] * fy[:, None]
    return out.astype(np.float32)


def resize_and_binarize(mask_f32: np.ndarray, original_wh: Tuple[int, int], threshold: float = 0.5) -> np.ndarray:
    """reference src/predict.py:620-625: bilinear resize of the float output to the original size, > thr -> {0,255}."""
    m = resize_linear_f32(mask_f32, int(original_wh[0]), int(original_wh[1]))
    return (m > np.float32(threshold)).astype(np.uint8) * 255


# ------------------------------------------------------------------------------------------------------------
# structuring elements and binary morphology (cv2.getStructuringElement / erode / dilate / morphologyEx)
# ------------------------------------------------------------------------------------------------------------
MORPH_RECT, MORPH_CROSS, MORPH_ELLIPSE = 0, 1, 2


def structuring_element(shape: int, ksize: Tuple[int, int]) -> np.ndarray:
    """cv2.getStructuringElement(shape, (cols, rows)) with the default anchor."""
    cols, rows = ksize
    el = np.zeros((rows, cols), np.uint8)
    if shape == MORPH_RECT or rows == 1 and cols == 1:
        el[:] = 1
        return el
    r, c = rows // 2, cols // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    for i in range(rows):
        if shape == MORPH_CROSS:
            j1, j2 = (0, cols) if i == r else (c, c + 1)
        else:
            dy = i - r
            if abs(dy) <= r:
                dx = int(round(c * math.sqrt((r * r - dy * dy) * inv_r2)))       # saturate_cast<int>(double)
                j1, j2 = max(c - dx, 0), min(c + dx + 1, cols)
            else:
                j1 = j2 = 0
        el[i, j1:j2] = 1
    return el


def _morph_once(img: np.ndarray, el: np.ndarray, dilate: bool) -> np.ndarray:
    """One erosion / dilation with the anchor at (cols//2, rows//2) and OpenCV's default constant border
    (out-of-image pixels never win: +inf for erosion, -inf for dilation):
        dst(y,x) = min|max over el(i,j) != 0 of src(y + i - ay, x + j - ax)."""
    rows, cols = el.shape
    ay, ax = rows // 2, cols // 2
    h, w = img.shape
    fill = 0 if dilate else 255
    pad = np.full((h + rows, w + cols), fill, np.uint8)
    pad[ay:ay + h, ax:ax + w] = img
    out = np.full((h, w), fill, np.uint8)
    for i in range(rows):
        for j in range(cols):
            if el[i, j]:
                win = pad[i:i + h, j:j + w]
                out = np.maximum(out, win) if dilate else np.minimum(out, win)
    return out


def erode(img, el, iterations=1):
    for _ in range(iterations):
        img = _morph_once(img, el, False)
    return img


def dilate(img, el, iterations=1):
    for _ in range(iterations):
        img = _morph_once(img, el, True)
    return img


def morph_open(img, el, iterations=1):
    return dilate(erode(img, el, iterations), el, iterations)


def morph_close(img, el, iterations=1):
    return erode(dilate(img, el, iterations), el, iterations)


# ------------------------------------------------------------------------------------------------------------
# 8-connected components with OpenCV's label order
# ------------------------------------------------------------------------------------------------------------
def connected_components_8(mask: np.ndarray):
    """labels (int32, 0 = background) and stats [n, 5] = (left, top, width, height, area) like
    cv2.connectedComponentsWithStats(mask, connectivity=8).

    Label ORDER: OpenCV's 8-connectivity algorithms scan 2x2 blocks in raster order and number components by the
    first block that opens them, so component a precedes component b iff a's first block (min over its pixels of
    (y // 2, x // 2) in raster order) comes first.  For almost every image that equals first-pixel raster order; it
    differs when a component's first pixel sits in the second row of an earlier block row."""
    h, w = mask.shape
    fg = mask != 0
    parent = np.arange(h * w, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    ys, xs = np.nonzero(fg)
    for y, x in zip(ys.tolist(), xs.tolist()):
        p = y * w + x
        for dy, dx in ((0, -1), (-1, -1), (-1, 0), (-1, 1)):
            yy, xx = y + dy, x + dx
            if 0 <= yy < h and 0 <= xx < w and fg[yy, xx]:
                ra, rb = find(p), find(yy * w + xx)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(y * w + x) for y, x in zip(ys.tolist(), xs.tolist())], dtype=np.int64)
    w2 = (w + 1) // 2
    key = (ys // 2) * w2 + (xs // 2)
    first = {}
    for r, k in zip(roots.tolist(), key.tolist()):
        if r not in first or k < first[r]:
            first[r] = k
    order = sorted(first, key=lambda r: first[r])
    lab_of = {r: i + 1 for i, r in enumerate(order)}
    labels = np.zeros((h, w), np.int32)
    labels[ys, xs] = [lab_of[r] for r in roots.tolist()]
    n = len(order) + 1
    stats = np.zeros((n, 5), np.int32)
    stats[0] = (0, 0, w, h, h * w - len(ys)) if len(ys) < h * w else (0, 0, 0, 0, 0)
    if len(ys) < h * w:
        by, bx = np.nonzero(~fg)
        stats[0] = (bx.min(), by.min(), bx.max() - bx.min() + 1, by.max() - by.min() + 1, len(by))
    for i in range(1, n):
        py, px = np.nonzero(labels == i)
        stats[i] = (px.min(), py.min(), px.max() - px.min() + 1, py.max() - py.min() + 1, len(py))
    return labels, stats


# ------------------------------------------------------------------------------------------------------------
# reference src/predict.py:161-301  _optimize_mask family
# ------------------------------------------------------------------------------------------------------------
def _binarize_127(mask):
    return np.where(mask > 127, 255, 0).astype(np.uint8)             # cv2.threshold(mask, 127, 255, THRESH_BINARY)


def _gauss3_threshold(mask):
    """cv2.GaussianBlur(mask, (3,3), 0.5) then threshold 127 (reference :270-271).  On a {0,255} image this is the
    identity: the centre weight of the 3x3 kernel is 0.619 (> 127/255) and all other weights sum to 0.381 (< 127/255),
    whatever the border mode and the fixed-point rounding of the 8-bit path - pinned against cv2 in the tests."""
    return _binarize_127(mask)


def optimize_watermark_mask(mask):
    """reference src/predict.py:232-273."""
    E = lambda k: structuring_element(MORPH_ELLIPSE, (k, k))      # noqa: E731
    mask = morph_open(mask, E(3), 1)
    mask = morph_close(mask, E(7), 3)
    mask = morph_close(mask, E(11), 2)
    mask = dilate(mask, E(9), 2)
    labels, stats = connected_components_8(mask)
    if stats.shape[0] > 1:
        areas = stats[1:, 4]
        largest = 1 + int(np.argmax(areas))                       # first maximum in label order
        mask = (labels == largest).astype(np.uint8) * 255
        if stats[largest, 4] < 500:
            mask = np.zeros_like(mask)
            for i in range(1, stats.shape[0]):
                if stats[i, 4] > 200:
                    mask[labels == i] = 255
    return _gauss3_threshold(mask)


def optimize_text_mask(mask):
    """reference src/predict.py:192-230."""
    E = lambda k: structuring_element(MORPH_ELLIPSE, (k, k))      # noqa: E731
    mask = morph_open(mask, E(2), 1)
    mask = morph_close(mask, E(3), 2)
    mh = morph_close(mask, structuring_element(MORPH_RECT, (5, 1)), 1)
    mv = morph_close(mask, structuring_element(MORPH_RECT, (1, 5)), 1)
    mask = mh | mv
    mask = dilate(mask, E(4), 1)
    return _keep_area_above(mask, 50)


def optimize_mixed_mask(mask):
    """reference src/predict.py:275-301."""
    E = lambda k: structuring_element(MORPH_ELLIPSE, (k, k))      # noqa: E731
    mask = morph_open(mask, E(2), 1)
    mask = morph_close(mask, E(5), 2)
    mask = dilate(mask, E(6), 1)
    return _keep_area_above(mask, 100)


def _keep_area_above(mask, thr):
    labels, stats = connected_components_8(mask)
    if stats.shape[0] > 1:
        out = np.zeros_like(mask)
        for i in range(1, stats.shape[0]):
            if stats[i, 4] > thr:
                out[labels == i] = 255
        return out
    return mask


def optimize_mask(mask, mask_type="watermark"):
    """reference src/predict.py:161-190."""
    if mask is None:
        return mask
    mask = _binarize_127(mask)
    if mask_type == "text":
        return optimize_text_mask(mask)
    if mask_type == "mixed":
        return optimize_mixed_mask(mask)
    return optimize_watermark_mask(mask)


def analyze_text_features(mask_binary) -> float:
    """reference src/predict.py:443-508 (geometric text score of the raw binary mask)."""
    if mask_binary is None or np.sum(mask_binary) == 0:
        return 0.0
    _, stats = connected_components_8(mask_binary)
    n = stats.shape[0]
    if n <= 1:
        return 0.0
    indicators, total = 0, n - 1
    for i in range(1, n):
        area, width, height = int(stats[i, 4]), int(stats[i, 2]), int(stats[i, 3])
        if area == 0 or width == 0 or height == 0:
            continue
        aspect = max(width, height) / min(width, height)
        density = area / (width * height)
        score = 0
        if 1 <= aspect <= 5:
            score += 0.3
        elif 5 < aspect <= 10:
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
            indicators += 1
    ratio = indicators / total
    if total >= 3 and ratio > 0.5:
        return min(ratio + 0.2, 1.0)
    return ratio
