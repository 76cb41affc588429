"""Host side of the image kernels either side of the network: ragged batches.

Thin ctypes wrappers over ``uwm_resize_bilinear_u8`` / ``uwm_mask_upscale_threshold`` / ``uwm_mask_postprocess`` /
``uwm_mask_text_features`` (include/uwm.h) on torch-owned CUDA buffers.  Images of different sizes travel in ONE
packed uint8 buffer; a :class:`RaggedBatch` carries the descriptor table (offset, width, height, pitch) on the
host (ctypes array) and on the device (an int64-viewed uint8 tensor).

Mirrors, in the reference (src/predict.py):
  * :588-602  cv2.imread -> cvtColor -> A.Resize(S,S)                  -> :func:`resize_bilinear_u8` (swap_rb)
  * :620-625  cv2.resize(mask, (W0,H0)); (mask > thr) * 255           -> :func:`mask_upscale_threshold`
  * :161-301  _optimize_mask / _optimize_{watermark,text,mixed}_mask  -> :func:`mask_postprocess`
  * :443-508  _analyze_text_features                                  -> :func:`mask_text_features`
There is no CPU fallback: every function raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MODE_WATERMARK, MODE_TEXT, MODE_MIXED = 0, 1, 2
MODES = {"watermark": MODE_WATERMARK, "text": MODE_TEXT, "mixed": MODE_MIXED}
OP_ERODE, OP_DILATE, OP_OPEN, OP_CLOSE = 0, 1, 2, 3
SHAPE_RECT, SHAPE_CROSS, SHAPE_ELLIPSE = 0, 1, 2        # cv2.MORPH_RECT / CROSS / ELLIPSE


class RaggedBatch:
    """Descriptor table of n images inside one packed buffer (elements = bytes for uint8 buffers)."""

    def __init__(self, sizes_wh: Sequence[Tuple[int, int]], channels: int = 1, device=None, align: int = 16,
                 pitches: Optional[Sequence[int]] = None, offsets: Optional[Sequence[int]] = None):
        self.n = len(sizes_wh)
        self.sizes = [(int(w), int(h)) for w, h in sizes_wh]
        self.channels = channels
        self.host = (_lib.ImageDesc * self.n)()
        off = 0
        for i, (w, h) in enumerate(self.sizes):
            pitch = int(pitches[i]) if pitches is not None else w * channels
            o = int(offsets[i]) if offsets is not None else off
            self.host[i] = _lib.ImageDesc(o, w, h, pitch, 0)
            off = max(off, o + pitch * h)
            off = (off + align - 1) // align * align
        self.total = off
        raw = np.frombuffer(bytes(self.host), dtype=np.uint8)          # copies the table
        self._pinned = torch.from_numpy(raw.copy())
        self.device_table = None
        if device is not None:
            self.to(device)

    def to(self, device, stream_non_blocking: bool = False):
        self.device_table = self._pinned.to(device, non_blocking=stream_non_blocking)
        return self

    @property
    def h_ptr(self):
        return C.cast(self.host, C.c_void_p)

    @property
    def d_ptr(self):
        if self.device_table is None:
            raise RuntimeError("RaggedBatch.to(device) first")
        return self.device_table.data_ptr()

    def view(self, packed: torch.Tensor, i: int) -> torch.Tensor:
        """Image i of a packed buffer as an [h, w(, c)] strided view."""
        d = self.host[i]
        rows = packed[d.offset:d.offset + d.pitch * d.height].view(d.height, d.pitch)
        v = rows[:, :d.width * self.channels]
        return v.view(d.height, d.width, self.channels) if self.channels > 1 else v


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (the B200 path has no CPU fallback)")


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def resize_bilinear_u8(packed: torch.Tensor, batch: RaggedBatch, dst_w: int, dst_h: int, swap_rb: bool = False,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR)`` for every uint8 HxWx3 image of the packed
    buffer -> uint8 [n, dst_h, dst_w, 3] (bit-exact with OpenCV).  ``swap_rb``: BGR in, RGB out."""
    _need_cuda(packed, "resize_bilinear_u8")
    if packed.dtype != torch.uint8 or batch.channels != 3:
        raise TypeError("resize_bilinear_u8: packed uint8 buffer of 3-channel images expected")
    if out is None:
        out = torch.empty(batch.n, dst_h, dst_w, 3, dtype=torch.uint8, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(_lib.load().uwm_resize_bilinear_u8(packed.data_ptr(), batch.h_ptr, batch.d_ptr, batch.n, dst_w, dst_h,
                                                      int(swap_rb), out.data_ptr(), _stream(packed)),
                   "uwm_resize_bilinear_u8")
    return out


def mask_upscale_threshold(maps: torch.Tensor, batch: RaggedBatch, threshold: float, out: Optional[torch.Tensor] = None,
                           return_float: bool = False):
    """``(cv2.resize(maps[i], (W0_i, H0_i)) > threshold) * 255`` for every image -> packed uint8 masks laid out as
    ``batch`` describes.  ``maps``: fp32 [n, h, w] or [n, 1, h, w].  ``return_float`` also returns the resized float
    maps (parity tap, same packed layout in floats)."""
    _need_cuda(maps, "mask_upscale_threshold")
    if maps.dim() == 4:
        maps = maps[:, 0]
    maps = maps.contiguous().float()
    n, sh, sw = maps.shape
    if n != batch.n:
        raise ValueError("mask_upscale_threshold: batch size mismatch")
    if out is None:
        out = torch.empty(batch.total, dtype=torch.uint8, device=maps.device)
    fl = torch.empty(batch.total, dtype=torch.float32, device=maps.device) if return_float else None
    with torch.cuda.device(maps.device):
        _lib.check(_lib.load().uwm_mask_upscale_threshold(maps.data_ptr(), n, sw, sh, batch.h_ptr, batch.d_ptr,
                                                          float(threshold), out.data_ptr(),
                                                          fl.data_ptr() if fl is not None else None, _stream(maps)),
                   "uwm_mask_upscale_threshold")
    return (out, fl) if return_float else out


def _workspace(batch: RaggedBatch, device) -> torch.Tensor:
    nbytes = int(_lib.load().uwm_mask_postprocess_workspace(batch.h_ptr, batch.n))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def mask_postprocess(masks: torch.Tensor, batch: RaggedBatch, mode="watermark", workspace: Optional[torch.Tensor] = None):
    """reference ``_optimize_mask(mask, mask_type)`` on every mask of the packed buffer, in place (bit-exact)."""
    _need_cuda(masks, "mask_postprocess")
    m = MODES[mode] if isinstance(mode, str) else int(mode)
    ws = workspace if workspace is not None else _workspace(batch, masks.device)
    with torch.cuda.device(masks.device):
        _lib.check(_lib.load().uwm_mask_postprocess(masks.data_ptr(), batch.h_ptr, batch.d_ptr, batch.n, m, ws.data_ptr(),
                                                    ws.numel(), _stream(masks)), "uwm_mask_postprocess")
    return masks


def _score_mask() -> int:
    """Which (aspect, density, area) class triples score > 0.5 in the reference's float arithmetic
    (reference src/predict.py:476-497: score = 0; += 0.3|0.1; += 0.3|0.1; += 0.4|0.2; score > 0.5)."""
    bits = 0
    for ia, a in enumerate((0.3, 0.1, None)):
        for ib, b in enumerate((0.3, 0.1, None)):
            for ic, c in enumerate((0.4, 0.2, None)):
                score = 0
                if a is not None:
                    score += a
                if b is not None:
                    score += b
                if c is not None:
                    score += c
                if score > 0.5:
                    bits |= 1 << ((3 * ia + ib) * 3 + ic)
    return bits


SCORE_MASK = _score_mask()


def mask_text_features(masks: torch.Tensor, batch: RaggedBatch, workspace: Optional[torch.Tensor] = None) -> List[float]:
    """reference ``_analyze_text_features(mask_binary)`` per image: the geometric text score in [0, 1]."""
    _need_cuda(masks, "mask_text_features")
    ws = workspace if workspace is not None else _workspace(batch, masks.device)
    out = torch.empty(batch.n, 2, dtype=torch.int32, device=masks.device)
    with torch.cuda.device(masks.device):
        _lib.check(_lib.load().uwm_mask_text_features(masks.data_ptr(), batch.h_ptr, batch.d_ptr, batch.n, SCORE_MASK,
                                                      out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(masks)),
                   "uwm_mask_text_features")
    scores = []
    for ind, total in out.cpu().tolist():
        if total == 0:
            scores.append(0.0)
            continue
        ratio = ind / total
        scores.append(min(ratio + 0.2, 1.0) if (total >= 3 and ratio > 0.5) else ratio)     # reference :499-508
    return scores


def mask_morphology(masks: torch.Tensor, batch: RaggedBatch, op: int, shape: int, ksize: Tuple[int, int],
                    iterations: int = 1):
    """One cv2 morphology call on every mask, in place (parity-test entry)."""
    _need_cuda(masks, "mask_morphology")
    ws = _workspace(batch, masks.device)
    with torch.cuda.device(masks.device):
        _lib.check(_lib.load().uwm_mask_morphology(masks.data_ptr(), batch.h_ptr, batch.d_ptr, batch.n, op, shape,
                                                   int(ksize[0]), int(ksize[1]), iterations, ws.data_ptr(), ws.numel(),
                                                   _stream(masks)), "uwm_mask_morphology")
    return masks


def mask_components(masks: torch.Tensor, batch: RaggedBatch, want_bbox: bool = True):
    """8-connected components of every mask (parity-test entry): per-pixel root index (-1 background) and, at each
    root pixel, area / OpenCV label-order key / bbox.  Arrays are packed per image at pixel offsets (no pitch)."""
    _need_cuda(masks, "mask_components")
    npx = sum(w * h for w, h in batch.sizes)
    dev = masks.device
    labels = torch.empty(npx, dtype=torch.int32, device=dev)
    area = torch.empty(npx, dtype=torch.int32, device=dev)
    order = torch.empty(npx, dtype=torch.int32, device=dev)
    bbox = torch.empty(npx, 4, dtype=torch.int32, device=dev) if want_bbox else None
    ws = _workspace(batch, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().uwm_mask_components(masks.data_ptr(), batch.h_ptr, batch.d_ptr, batch.n, labels.data_ptr(),
                                                   area.data_ptr(), order.data_ptr(),
                                                   bbox.data_ptr() if bbox is not None else None, ws.data_ptr(),
                                                   ws.numel(), _stream(masks)), "uwm_mask_components")
    return labels, area, order, bbox


def mask_component_summary(masks: torch.Tensor, batch: RaggedBatch) -> List[Tuple[int, int, int]]:
    """Per mask: (foreground pixels, 8-connected components, largest component area) - the statistics the reference's
    model_selector takes from ``cv2.connectedComponentsWithStats`` (src/scripts/model_selector.py:171-197)."""
    _need_cuda(masks, "mask_component_summary")
    ws = _workspace(batch, masks.device)
    out = torch.empty(batch.n, 3, dtype=torch.int32, device=masks.device)
    with torch.cuda.device(masks.device):
        _lib.check(_lib.load().uwm_mask_component_summary(masks.data_ptr(), batch.h_ptr, batch.d_ptr, batch.n, out.data_ptr(),
                                                          ws.data_ptr(), ws.numel(), _stream(masks)),
                   "uwm_mask_component_summary")
    return [tuple(r) for r in out.cpu().tolist()]
