"""Mask-inference surface of the reference predictor, batched and B200-native.

Mirrors the UNet step of ``WatermarkPredictor`` (reference src/predict.py): model load (:68-99), file discovery /
skip-existing (:114-160), single-image ``predict_mask`` (:303-368), ``step1_batch_predict_watermark_masks``
(:560-664) with its watermark-type detection (:414-558) and mask post-processing (:161-301).  The IOPaint / OCR
steps 2-5 of the repair pipeline are out of scope.

Where the work runs (per batch of images of arbitrary sizes):

    CPU threads   cv2.imread (BGR, original size) -> one pinned, packed staging buffer          reference :591-595
    GPU           resize_u8 (cv2 INTER_LINEAR, bit-exact; BGR->RGB folded in)                   reference :598-602
                  -> prep (Normalize) + Unet forward (tcgen05 kernels)                          reference :605-617
                  -> bilinear upscale to each original size + threshold -> uint8 masks          reference :620-625
                  -> [_analyze_text_features on the GPU; Canny/Sobel statistics on CPU threads] reference :414-558
                  -> _optimize_mask (bit-packed morphology, 8-connected components, area rules)  reference :161-301
    CPU threads   cv2.imwrite(<stem>_mask.png), watermark ratio                                 reference :629-639

Mask conventions, selectable with ``sigmoid``:
  * ``sigmoid=False`` (default, the reference predictor): ``cv2.resize(out) > thr``   — predict.py:620-625
  * ``sigmoid=True``:  ``cv2.resize(sigmoid(out)) > thr``                            — watermark_filter.py:136-150
Post-processing (``post_process=True`` = the reference's step 1): ``mask_type='auto'`` detects the watermark type per
image like the reference; a fixed type ('watermark' | 'text' | 'mixed') skips the detection (as ``predict_mask``
does, reference :303-368).
"""
from __future__ import annotations

import glob
import logging
import os
import random
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import imgproc
from .config import get_cfg_defaults, install_yacs_shim, update_config
from .unet_model import create_model_from_config

logger = logging.getLogger(__name__)

IMAGE_EXTENSIONS = ["*.jpg", "*.jpeg", "*.png", "*.bmp", "*.tiff", "*.webp"]


def shard_for_rank(items: Sequence, rank: int, world_size: int) -> List:
    """Data-parallel partition of the sorted work list: rank r takes items r, r+W, r+2W, ...
    (no collective on the inference path)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(items[rank::world_size])


def load_checkpoint_state(model_path: str, map_location="cpu") -> Tuple[Dict[str, torch.Tensor], dict]:
    """``torch.load`` a reference checkpoint: either the trainer's dict
    ``{'epoch','model_state_dict','val_loss','val_metrics','config'}`` (reference src/train.py:428-435) or a
    bare state dict (reference src/predict.py:80-91)."""
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"模型文件不存在: {model_path}")
    install_yacs_shim()
    ckpt = torch.load(model_path, map_location=map_location, weights_only=False)
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
        info = {"epoch": ckpt.get("epoch", "Unknown"), "val_loss": ckpt.get("val_loss", "Unknown"),
                "val_metrics": ckpt.get("val_metrics", {})}
        return ckpt["model_state_dict"], info
    return ckpt, {"epoch": "Unknown", "val_loss": "Unknown"}


def analyze_ocr_features(image_rgb: np.ndarray, mask_binary: np.ndarray) -> float:
    """Edge-density / gradient-direction score of the masked image (reference src/predict.py:510-558, same OpenCV
    calls).  Canny + float64 Sobel statistics: host work, run on the decode threads."""
    import cv2
    try:
        masked = cv2.bitwise_and(image_rgb, cv2.cvtColor(mask_binary, cv2.COLOR_GRAY2RGB))
        gray = cv2.cvtColor(masked, cv2.COLOR_RGB2GRAY)
        n_mask = np.sum(mask_binary > 0)
        edge_density = np.sum(cv2.Canny(gray, 50, 150) > 0) / n_mask if n_mask > 0 else 0
        angles = np.arctan2(cv2.Sobel(gray, cv2.CV_64F, 0, 1, ksize=3), cv2.Sobel(gray, cv2.CV_64F, 1, 0, ksize=3))
        angle_variance = np.var(angles[mask_binary > 0]) if n_mask > 0 else 0
        score = 0
        if 0.1 <= edge_density <= 0.4:
            score += 0.5
        elif 0.05 <= edge_density < 0.1 or 0.4 < edge_density <= 0.6:
            score += 0.2
        if 1.0 <= angle_variance <= 3.0:
            score += 0.5
        elif 0.5 <= angle_variance < 1.0 or 3.0 < angle_variance <= 4.0:
            score += 0.2
        return min(score, 1.0)
    except Exception as e:  # noqa: BLE001 - reference :556-558
        logger.debug("OCR特征分析失败: %s", e)
        return 0.0


def watermark_type_from_scores(text_score: float, ocr_score: float) -> str:
    """reference src/predict.py:430-438."""
    total = text_score * 0.6 + ocr_score * 0.4
    if total > 0.7:
        return "text"
    if total > 0.3:
        return "mixed"
    return "watermark"


def enhance_text_features(image_rgb: np.ndarray) -> np.ndarray:
    """reference src/predict.py:370-404 (CLAHE / Canny / sharpen pre-enhancement of the 'text' and 'mixed' single-image
    modes), the same OpenCV calls on the host."""
    import cv2
    gray = cv2.cvtColor(image_rgb, cv2.COLOR_RGB2GRAY)
    enhanced_gray = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(gray)
    edges = cv2.dilate(cv2.Canny(enhanced_gray, 50, 150), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2)), iterations=1)
    out = image_rgb.copy()
    edge_mask = edges > 0
    for i in range(3):
        ch = out[:, :, i].astype(np.float32)
        ch[edge_mask] = np.clip(ch[edge_mask] * 1.2, 0, 255)
        out[:, :, i] = ch.astype(np.uint8)
    return cv2.filter2D(out, -1, np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]]))


class _Staging:
    """Pinned host staging buffers that grow on demand (two of each: batch i+1 is filled while batch i is in flight)."""

    def __init__(self):
        self.bufs: Dict[Tuple[str, int], torch.Tensor] = {}

    def get(self, kind: str, slot: int, nbytes: int) -> torch.Tensor:
        b = self.bufs.get((kind, slot))
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory()
            self.bufs[(kind, slot)] = b
        return b


class WatermarkPredictor:
    """UNet mask predictor (step 1 of the reference's pipeline), batched, pre/post-processing on the GPU."""

    def __init__(self, model_path, config_path=None, config=None, device="cuda", batch_size: Optional[int] = None,
                 sigmoid: bool = False, num_workers: int = 8, post_process: Optional[bool] = None, mask_type: str = "auto"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("unet_watermark_b200 runs the mask path on CUDA (sm_100a) only; "
                               f"device={device!r} is not supported (no CPU fallback)")
        if config is not None:
            self.cfg = config
        else:
            self.cfg = get_cfg_defaults()
            if config_path and os.path.exists(config_path):
                update_config(self.cfg, config_path)
        if mask_type not in ("auto", "watermark", "text", "mixed"):
            raise ValueError(f"mask_type {mask_type!r}: expected auto | watermark | text | mixed")
        self.sigmoid = sigmoid
        self.mask_type = mask_type
        self.post_process = bool(getattr(self.cfg.PREDICT, "POST_PROCESS", True)) if post_process is None else bool(post_process)
        self.img_size = int(self.cfg.DATA.IMG_SIZE)
        self.batch_size = int(batch_size or getattr(self.cfg.PREDICT, "BATCH_SIZE", 8))
        self.num_workers = num_workers
        self._staging = _Staging()
        self.model, self.model_info = self._load_unet_model(model_path)
        logger.info("mask predictor ready on %s (batch %d, %dx%d)", self.device, self.batch_size, self.img_size,
                    self.img_size)

    # -- model -----------------------------------------------------------------------------------
    def _load_unet_model(self, model_path):
        if not os.path.exists(model_path):
            raise FileNotFoundError(f"模型文件不存在: {model_path}")
        try:
            model = create_model_from_config(self.cfg).to(self.device)
            state, info = load_checkpoint_state(model_path, map_location="cpu")
            model.load_state_dict(state)
            model.eval()
            return model, info
        except Exception:
            torch.cuda.empty_cache()
            raise

    # -- files -----------------------------------------------------------------------------------
    def _get_image_files(self, input_folder, output_folder=None, limit=None, rank: int = 0, world_size: int = 1,
                         seed: int = 0):
        """Same discovery rules as the reference (:114-160): 6 extensions x 2 cases, de-duplicated and sorted,
        images whose ``<stem>_mask.png`` exists are skipped, optional random ``limit``.

        Multi-rank order of operations (a stable partition needs it): the FULL sorted list is sharded first
        (rank r owns files r, r+W, ...), the skip-existing filter then runs inside each shard - a rank that lists
        the directory late, after other ranks have written masks, still owns exactly the same files - and
        ``limit`` selects the same seeded random subset of the unfiltered list on every rank before sharding
        (the reference shuffles unseeded in its single process, :155-157)."""
        files = []
        for ext in IMAGE_EXTENSIONS:
            files.extend(glob.glob(os.path.join(input_folder, ext)))
            files.extend(glob.glob(os.path.join(input_folder, ext.upper())))
        files = sorted(set(files))
        if limit is not None and limit > 0 and len(files) > limit:
            if world_size > 1:
                random.Random(seed).shuffle(files)
                files = sorted(files[:limit])
            else:                                          # single process: the reference's order (filter, then shuffle)
                files = self._skip_existing(files, output_folder)
                random.shuffle(files)
                return files[:limit]
        files = shard_for_rank(files, rank, world_size)
        return self._skip_existing(files, output_folder)

    @staticmethod
    def _skip_existing(files, output_folder):
        if not (output_folder and os.path.exists(output_folder)):
            return files
        return [p for p in files
                if not os.path.exists(os.path.join(output_folder, f"{os.path.splitext(os.path.basename(p))[0]}_mask.png"))]

    # -- decode (CPU threads) ----------------------------------------------------------------------
    @staticmethod
    def _decode(path: str):
        """cv2.imread: BGR uint8 at the original size (reference :591); None if unreadable."""
        import cv2
        return cv2.imread(path)

    # -- one batch on the GPU ----------------------------------------------------------------------
    def _masks_for_images(self, images_bgr: List[np.ndarray], threshold: float, slot: int = 0,
                          mask_types: Optional[List[str]] = None, pool: Optional[ThreadPoolExecutor] = None,
                          morphology: Optional[Sequence[Tuple[int, int, Tuple[int, int], int]]] = None,
                          sigmoid: Optional[bool] = None) -> Tuple[List[np.ndarray], List[str]]:
        """BGR uint8 images of arbitrary sizes -> (uint8 {0,255} masks at the original sizes, mask types used).
        ``morphology``: a list of (op, shape, ksize, iterations) cv2-morphology steps run on the GPU INSTEAD of the
        reference predictor's ``_optimize_mask`` (the other callers of the seam post-process differently, e.g.
        reference src/scripts/watermark_filter.py:159-166); ``sigmoid`` overrides the predictor's convention."""
        n = len(images_bgr)
        s = self.img_size
        dev = self.device
        sizes = [(im.shape[1], im.shape[0]) for im in images_bgr]
        # 1. pack into pinned staging, one H2D copy
        src = imgproc.RaggedBatch(sizes, channels=3)
        stage = self._staging.get("in", slot, src.total)
        host_np = stage.numpy()

        def put(i):
            d = src.host[i]
            host_np[d.offset:d.offset + d.pitch * d.height].reshape(d.height, d.pitch)[:] = images_bgr[i].reshape(d.height, -1)
        if pool is not None and n > 1:
            list(pool.map(put, range(n)))
        else:
            for i in range(n):
                put(i)
        with torch.cuda.device(dev):
            packed = stage[:src.total].to(dev, non_blocking=True)
            src.to(dev, stream_non_blocking=True)
            # 2. resize to the network size (bit-exact cv2 INTER_LINEAR; BGR -> RGB in the read), forward
            x = imgproc.resize_bilinear_u8(packed, src, s, s, swap_rb=True)
            use_sigmoid = self.sigmoid if sigmoid is None else sigmoid
            maps = self.model.predict_proba(x) if use_sigmoid else self.model(x)           # fp32 [n,1,S,S]
            # 3. bilinear upscale to every original size + threshold -> packed uint8 masks
            dst = imgproc.RaggedBatch(sizes, channels=1, device=dev)
            masks = imgproc.mask_upscale_threshold(maps, dst, threshold)
            types = list(mask_types) if mask_types is not None else [self.mask_type] * n
            if morphology is not None:
                for op, shape, ksize, iters in morphology:
                    imgproc.mask_morphology(masks, dst, op, shape, ksize, iters)
            elif self.post_process:
                ws = None
                if any(t == "auto" for t in types):
                    # reference _detect_watermark_type: geometric score on the GPU, Canny/Sobel statistics on CPU threads
                    text_scores = imgproc.mask_text_features(masks, dst)
                    raw = masks.cpu()
                    def ocr(i):
                        import cv2
                        return analyze_ocr_features(cv2.cvtColor(images_bgr[i], cv2.COLOR_BGR2RGB), dst.view(raw, i).numpy())
                    idx = [i for i, t in enumerate(types) if t == "auto"]
                    ocr_scores = list(pool.map(ocr, idx)) if pool is not None else [ocr(i) for i in idx]
                    for i, o in zip(idx, ocr_scores):
                        types[i] = watermark_type_from_scores(text_scores[i], o)
                # 4. _optimize_mask per type (one ragged launch sequence per type present in the batch)
                for t in sorted(set(types)):
                    sel = [i for i, ti in enumerate(types) if ti == t]
                    if len(sel) == n:
                        imgproc.mask_postprocess(masks, dst, t)
                    else:
                        sub = imgproc.RaggedBatch([sizes[i] for i in sel], channels=1, device=dev,
                                                  pitches=[dst.host[i].pitch for i in sel],
                                                  offsets=[dst.host[i].offset for i in sel])
                        imgproc.mask_postprocess(masks, sub, t)
            out_stage = self._staging.get("out", slot, dst.total)
            out_stage[:dst.total].copy_(masks, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        out_np = out_stage.numpy()
        result = []
        for i in range(n):
            d = dst.host[i]
            result.append(out_np[d.offset:d.offset + d.pitch * d.height].reshape(d.height, d.pitch)[:, :d.width].copy())
        return result, types

    # -- public API ------------------------------------------------------------------------------
    def predict_mask(self, image_path, mask_type="watermark"):
        """Single image -> uint8 {0,255} mask at the original resolution (reference :303-368): forward, bilinear
        resize to the original size, threshold, ``_optimize_mask(mask, mask_type)``."""
        import cv2
        image = self._decode(image_path)
        if image is None:
            raise ValueError(f"无法读取图像: {image_path}")
        if mask_type in ("text", "mixed"):                      # reference :323-325 (host-side pre-enhancement)
            image = cv2.cvtColor(enhance_text_features(cv2.cvtColor(image, cv2.COLOR_BGR2RGB)), cv2.COLOR_RGB2BGR)
        thr = float(getattr(self.cfg.PREDICT, "THRESHOLD", 0.5))
        masks, _ = self._masks_for_images([image], thr, mask_types=[mask_type])
        return masks[0]

    def step1_batch_predict_watermark_masks(self, input_folder, mask_output_folder, limit=None, rank: int = 0,
                                            world_size: int = 1):
        """Batched step 1 (reference :560-664): writes ``<stem>_mask.png`` for every unprocessed image and
        returns ``[{image_path, mask_path, watermark_ratio}]`` for images with a non-empty mask."""
        import cv2
        os.makedirs(mask_output_folder, exist_ok=True)
        files = self._get_image_files(input_folder, mask_output_folder, limit=limit, rank=rank, world_size=world_size)
        if not files:
            logger.warning("在 %s 中未找到未处理的图像文件", input_folder)
            return []
        thr = float(getattr(self.cfg.PREDICT, "THRESHOLD", 0.5))
        bs = self.batch_size
        processed: List[dict] = []

        def write(path, mask):
            base = os.path.splitext(os.path.basename(path))[0]
            mask_path = os.path.join(mask_output_folder, f"{base}_mask.png")
            cv2.imwrite(mask_path, mask)
            wm = int(np.count_nonzero(mask))
            if wm == 0:                                        # reference :642-645: all-black mask, not reported
                return None
            return {"image_path": path, "mask_path": mask_path,
                    "watermark_ratio": wm / float(mask.shape[0] * mask.shape[1])}

        with ThreadPoolExecutor(max_workers=self.num_workers) as pool:
            chunks = [files[i:i + bs] for i in range(0, len(files), bs)]
            pending = [pool.submit(self._decode, p) for p in chunks[0]]
            writes = []
            for ci, chunk in enumerate(chunks):
                images = [f.result() for f in pending]
                if ci + 1 < len(chunks):                      # decode the next batch while the GPU works on this one
                    pending = [pool.submit(self._decode, p) for p in chunks[ci + 1]]
                paths, imgs = [], []
                for path, im in zip(chunk, images):
                    if im is None:
                        logger.error("无法加载图像: %s", path)
                        continue
                    paths.append(path)
                    imgs.append(im)
                if not paths:
                    continue
                try:
                    masks, _ = self._masks_for_images(imgs, thr, slot=ci & 1, pool=pool)
                except Exception as e:  # noqa: BLE001 - per-batch failures are logged and skipped (:655-657)
                    logger.error("处理图像失败 %s: %s", paths, e)
                    continue
                writes.extend(pool.submit(write, p, m) for p, m in zip(paths, masks))      # PNG encode off the main thread
            for f in writes:
                r = f.result()
                if r is not None:
                    processed.append(r)
        return processed
