"""CUDA image kernels (csrc/uwm_imgproc.cu, through the C ABI) against cv2 - the dependency the reference calls -
against the numpy oracle and against the committed cv2 golden vectors.  Integer / byte work: BIT-EXACT.
Float resize: bit-exact with OpenCV's own algorithm (numpy oracle == cv2 with IPP off); against the IPP-routed
cv2.resize of the pip wheel |d| <= 5e-4 and masks agree on all but ~1e-5 of the pixels."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import imgproc_oracle as I
from tests import cv2_reference as R
from unet_watermark_b200 import imgproc as G

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imgproc_cv2.npz"))


def pack(images, channels, dev, pitch_pad=0):
    sizes = [(im.shape[1], im.shape[0]) for im in images]
    pitches = [im.shape[1] * channels + pitch_pad for im in images]
    rb = G.RaggedBatch(sizes, channels=channels, device=dev, pitches=pitches)
    buf = np.zeros(rb.total, np.uint8)
    for i, im in enumerate(images):
        d = rb.host[i]
        rows = buf[d.offset:d.offset + d.pitch * d.height].reshape(d.height, d.pitch)
        rows[:, :d.width * channels] = im.reshape(d.height, d.width * channels)
    return rb, torch.from_numpy(buf).to(dev)


def unpack(rb, packed, i):
    return rb.view(packed.cpu(), i).numpy().copy()


def test_resize_u8_ragged_batch_bit_exact_with_cv2(cuda_device):
    rng = np.random.default_rng(0)
    shapes = [(37, 53), (512, 512), (1024, 1024), (480, 640), (1080, 1920), (3, 5), (1, 1), (700, 333), (256, 256),
              (1024, 700), (2, 2000), (513, 511)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    rb, packed = pack(imgs, 3, cuda_device, pitch_pad=5)
    for (dw, dh) in ((512, 512), (64, 96)):
        out = G.resize_bilinear_u8(packed, rb, dw, dh).cpu().numpy()
        for i, im in enumerate(imgs):
            want = cv2.resize(im, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(out[i], want), (shapes[i], dw, dh)
            assert np.array_equal(out[i], I.resize_linear_u8(im, dw, dh))
    # BGR in (cv2.imread order), RGB out: cvtColor folded into the read
    out = G.resize_bilinear_u8(packed, rb, 128, 128, swap_rb=True).cpu().numpy()
    for i, im in enumerate(imgs):
        want = cv2.resize(cv2.cvtColor(im, cv2.COLOR_BGR2RGB), (128, 128), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(out[i], want)


def test_upscale_threshold_matches_opencv_algorithm(cuda_device):
    rng = np.random.default_rng(1)
    S = 128
    maps = rng.normal(0.4, 1.0, (7, S, S)).astype(np.float32)
    sizes = [(640, 480), (128, 128), (64, 64), (1000, 37), (91, 517), (256, 256), (130, 126)]      # (w, h)
    rb = G.RaggedBatch(sizes, channels=1, device=cuda_device)
    masks, fl = G.mask_upscale_threshold(torch.from_numpy(maps).to(cuda_device), rb, 0.5, return_float=True)
    masks, fl = masks.cpu(), fl.cpu()
    flips = total = 0
    for i, (w, h) in enumerate(sizes):
        d = rb.host[i]
        got_f = fl[d.offset:d.offset + d.pitch * h].view(h, d.pitch)[:, :w].numpy()
        got_m = rb.view(masks, i).numpy()
        want_f = I.resize_linear_f32(maps[i], w, h)
        assert np.array_equal(got_f, want_f), sizes[i]                 # OpenCV's own algorithm, bit for bit
        assert np.array_equal(got_m, I.resize_and_binarize(maps[i], (w, h), 0.5))
        ref = cv2.resize(maps[i], (w, h))                              # the pip wheel routes float32 through IPP
        assert np.abs(got_f - ref).max() <= 5e-4 * np.abs(maps[i]).max()
        flips += int((got_m != ((ref > 0.5).astype(np.uint8) * 255)).sum())
        total += got_m.size
    assert flips / total <= 1e-4, (flips, total)


@pytest.mark.parametrize("pitch_pad", [0, 3])
def test_morphology_bit_exact_with_cv2(cuda_device, pitch_pad):
    rng = np.random.default_rng(2)
    shapes = [(33, 47), (64, 64), (1, 100), (100, 1), (70, 200), (31, 32), (32, 33), (5, 5), (257, 129)]
    base = [(rng.random(s) < p).astype(np.uint8) * 255 for s, p in zip(shapes, [0.1, 0.5, 0.9, 0.5, 0.3, 0.7, 0.5, 0.5, 0.4])]
    base += [R.blob_mask(120, 333, 4, noise=0.02)]
    cases = [(G.OP_ERODE, cv2.MORPH_ELLIPSE, (3, 3), 1), (G.OP_DILATE, cv2.MORPH_ELLIPSE, (9, 9), 2),
             (G.OP_OPEN, cv2.MORPH_ELLIPSE, (2, 2), 1), (G.OP_CLOSE, cv2.MORPH_ELLIPSE, (7, 7), 3),
             (G.OP_CLOSE, cv2.MORPH_ELLIPSE, (11, 11), 2), (G.OP_CLOSE, cv2.MORPH_RECT, (5, 1), 1),
             (G.OP_CLOSE, cv2.MORPH_RECT, (1, 5), 1), (G.OP_DILATE, cv2.MORPH_ELLIPSE, (4, 4), 1),
             (G.OP_DILATE, cv2.MORPH_ELLIPSE, (6, 6), 1), (G.OP_ERODE, cv2.MORPH_CROSS, (5, 5), 2),
             (G.OP_OPEN, cv2.MORPH_RECT, (3, 3), 2), (G.OP_ERODE, cv2.MORPH_ELLIPSE, (13, 13), 1)]
    for op, shape, ks, iters in cases:
        rb, packed = pack(base, 1, cuda_device, pitch_pad)
        G.mask_morphology(packed, rb, op, shape, ks, iters)
        el = cv2.getStructuringElement(shape, ks)
        for i, m in enumerate(base):
            want = {G.OP_ERODE: lambda: cv2.erode(m, el, iterations=iters), G.OP_DILATE: lambda: cv2.dilate(m, el, iterations=iters),
                    G.OP_OPEN: lambda: cv2.morphologyEx(m, cv2.MORPH_OPEN, el, iterations=iters),
                    G.OP_CLOSE: lambda: cv2.morphologyEx(m, cv2.MORPH_CLOSE, el, iterations=iters)}[op]()
            assert np.array_equal(unpack(rb, packed, i), want), (op, shape, ks, iters, m.shape)


def _cv_labels_from_roots(labels, order, w, h):
    """Per-pixel root indices + per-root order keys -> OpenCV numbering (components sorted by order key)."""
    lab = labels.reshape(h, w)
    roots = np.unique(lab[lab >= 0])
    keys = order[roots]
    rank = {int(r): i + 1 for i, r in enumerate(roots[np.argsort(keys, kind="stable")])}
    out = np.zeros((h, w), np.int32)
    for r, k in rank.items():
        out[lab == r] = k
    return out, roots


def test_connected_components_equal_cv2_including_label_order(cuda_device):
    rng = np.random.default_rng(3)
    crafted = np.zeros((4, 6), np.uint8)
    crafted[1, 0] = 255; crafted[0, 3] = 255; crafted[3, 5] = 255; crafted[2, 2] = 255
    masks = [crafted, (rng.random((40, 40)) < 0.4).astype(np.uint8) * 255, (rng.random((64, 90)) < 0.55).astype(np.uint8) * 255,
             np.zeros((8, 8), np.uint8), np.full((9, 33), 255, np.uint8), R.blob_mask(300, 420, 6, noise=0.03),
             (rng.random((1, 70)) < 0.5).astype(np.uint8) * 255, (rng.random((513, 1)) < 0.5).astype(np.uint8) * 255]
    rb, packed = pack(masks, 1, cuda_device)
    labels, area, order, bbox = (t.cpu().numpy() for t in G.mask_components(packed, rb))
    off = 0
    for m in masks:
        h, w = m.shape
        n, want, stats, _ = cv2.connectedComponentsWithStats(m, connectivity=8)
        L, roots = _cv_labels_from_roots(labels[off:off + h * w], order[off:off + h * w], w, h)
        assert np.array_equal(L, want), m.shape
        for r in roots:
            k = L.reshape(-1)[r]
            assert area[off + r] == stats[k, cv2.CC_STAT_AREA]
            x0, y0, x1, y1 = bbox[off + r]
            assert (x0, y0, x1 - x0 + 1, y1 - y0 + 1) == tuple(stats[k, :4])
        off += h * w


@pytest.mark.parametrize("mode", ["watermark", "text", "mixed"])
def test_mask_postprocess_equals_reference_optimize_mask(cuda_device, mode):
    small = np.zeros((80, 80), np.uint8)
    small[5:8, 5:9] = 255; small[40:44, 40:45] = 255
    two = np.zeros((120, 120), np.uint8)                       # two components of equal area: cv2's label order decides
    two[60:90, 10:40] = 255; two[10:40, 70:100] = 255
    grey = (np.arange(64 * 64).reshape(64, 64) % 256).astype(np.uint8)
    masks = [R.blob_mask(96, 128, 0, noise=0.01), R.blob_mask(70, 200, 1, noise=0.03), R.blob_mask(333, 517, 2, noise=0.0),
             R.blob_mask(1080, 1920, 3, n_blobs=12, noise=0.001), small, two, grey, np.zeros((20, 30), np.uint8),
             np.full((31, 65), 255, np.uint8), GOLD["mask"]]
    rb, packed = pack(masks, 1, cuda_device, pitch_pad=7)
    G.mask_postprocess(packed, rb, mode)
    for i, m in enumerate(masks):
        got = unpack(rb, packed, i)
        assert np.array_equal(got, R.optimize_mask(m.copy(), mode)), (mode, i, m.shape)
    assert np.array_equal(unpack(rb, packed, len(masks) - 1), GOLD[f"opt_{mode}"])


def test_text_features_equal_reference(cuda_device):
    masks = [R.blob_mask(100, 140, s, n_blobs=3 + s, noise=0.01 * s) for s in range(6)] + [np.zeros((16, 16), np.uint8), GOLD["mask"]]
    rb, packed = pack(masks, 1, cuda_device)
    scores = G.mask_text_features(packed, rb)
    for m, s in zip(masks, scores):
        assert s == R.analyze_text_features(m)
    assert scores[-1] == float(GOLD["text_score"])


def test_golden_vectors_on_gpu(cuda_device):
    g = GOLD
    rb, packed = pack([g["u8_src"], g["u8_src_2x"]], 3, cuda_device)
    out = G.resize_bilinear_u8(packed, rb, 32, 32).cpu().numpy()
    assert np.array_equal(out[0], g["u8_to_32x32"]) and np.array_equal(out[1], g["u8_2x_to_32x32"])
    rb2 = G.RaggedBatch([(75, 41)], device=cuda_device)
    _, fl = G.mask_upscale_threshold(torch.from_numpy(g["f32_src"][None]).to(cuda_device), rb2, 0.5, return_float=True)
    assert np.array_equal(fl.cpu().numpy()[:75 * 41].reshape(41, 75), g["f32_to_75x41"])


def test_no_cpu_fallback():
    rb = G.RaggedBatch([(4, 4)], channels=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.resize_bilinear_u8(torch.zeros(48, dtype=torch.uint8), rb, 2, 2)
