"""Mask-inference surface of the reference predictor, batched and B200-native.

Mirrors the UNet step of ``WatermarkPredictor`` (reference src/predict.py): model load
(:68-99), file discovery / skip-existing (:114-160), single-image ``predict_mask`` (:303-368) and
``step1_batch_predict_watermark_masks`` (:560-664) — but runs the network through
``libuwm_b200.so`` in batches instead of one image per forward.  The IOPaint / OCR steps 2-5
and the OpenCV mask post-processing (``_optimize_mask``) are out of scope.

Mask conventions, selectable with ``sigmoid``:
  * ``sigmoid=False`` (default, the reference predictor): ``cv2.resize(out) > thr``   — predict.py:620-625
  * ``sigmoid=True``:  ``cv2.resize(sigmoid(out)) > thr``                            — watermark_filter.py:136-150
When the image already has the network size no resize happens and the uint8 mask comes straight
from the fused head kernel.
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


class WatermarkPredictor:
    """UNet mask predictor (step 1 of the reference's pipeline), batched."""

    def __init__(self, model_path, config_path=None, config=None, device="cuda", batch_size: Optional[int] = None,
                 sigmoid: bool = False, num_workers: int = 8):
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
        self.sigmoid = sigmoid
        self.img_size = int(self.cfg.DATA.IMG_SIZE)
        self.batch_size = int(batch_size or getattr(self.cfg.PREDICT, "BATCH_SIZE", 8))
        self.num_workers = num_workers
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

    # -- pre / post ------------------------------------------------------------------------------
    def _load_resized(self, path: str):
        """cv2.imread -> RGB -> bilinear resize to the network size (the Resize of get_val_transform,
        reference src/utils/dataset.py:389-395); Normalize is fused into the GPU prep kernel."""
        import cv2
        img = cv2.imread(path)
        if img is None:
            return None
        h0, w0 = img.shape[:2]
        rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        if (h0, w0) != (self.img_size, self.img_size):
            rgb = cv2.resize(rgb, (self.img_size, self.img_size), interpolation=cv2.INTER_LINEAR)
        return np.ascontiguousarray(rgb), (w0, h0)

    def _masks_for_batch(self, batch_u8: torch.Tensor, sizes: List[Tuple[int, int]], threshold: float):
        """uint8 NHWC host batch -> list of uint8 {0,255} masks at the original sizes."""
        import cv2
        s = self.img_size
        x = batch_u8.to(self.device, non_blocking=True)
        need_float = any(sz != (s, s) for sz in sizes)
        if not need_float:
            m = self.model.predict_mask(x, threshold, sigmoid=self.sigmoid).cpu().numpy()
            return [m[i] for i in range(len(sizes))]
        out = (self.model.predict_proba(x) if self.sigmoid else self.model(x))[:, 0].cpu().numpy()
        masks = []
        for i, (w0, h0) in enumerate(sizes):
            mi = out[i] if (w0, h0) == (s, s) else cv2.resize(out[i], (w0, h0))
            masks.append((mi > threshold).astype(np.uint8) * 255)
        return masks

    # -- public API ------------------------------------------------------------------------------
    def predict_mask(self, image_path, mask_type="watermark"):
        """Single image -> uint8 {0,255} mask at the original resolution (reference :303-368, without the
        OpenCV post-processing / text-enhancement branches, which are out of scope)."""
        item = self._load_resized(image_path)
        if item is None:
            raise ValueError(f"无法读取图像: {image_path}")
        rgb, size = item
        thr = float(getattr(self.cfg.PREDICT, "THRESHOLD", 0.5))
        batch = torch.from_numpy(rgb).unsqueeze(0)
        return self._masks_for_batch(batch, [size], thr)[0]

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
        s, bs = self.img_size, self.batch_size
        processed = []
        pinned = [torch.empty(bs, s, s, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
        with ThreadPoolExecutor(max_workers=self.num_workers) as pool:
            chunks = [files[i:i + bs] for i in range(0, len(files), bs)]
            pending = pool.map(self._load_resized, chunks[0]) if chunks else None
            for ci, chunk in enumerate(chunks):
                items = list(pending)
                if ci + 1 < len(chunks):                      # decode the next batch while the GPU works
                    pending = pool.map(self._load_resized, chunks[ci + 1])
                paths, sizes, buf = [], [], pinned[ci & 1]
                for path, item in zip(chunk, items):
                    if item is None:
                        logger.error("无法加载图像: %s", path)
                        continue
                    buf[len(paths)].copy_(torch.from_numpy(item[0]))
                    paths.append(path)
                    sizes.append(item[1])
                if not paths:
                    continue
                try:
                    masks = self._masks_for_batch(buf[:len(paths)], sizes, thr)
                except Exception as e:  # noqa: BLE001 - per-batch failures are logged and skipped (:655-657)
                    logger.error("处理图像失败 %s: %s", paths, e)
                    continue
                for path, mask in zip(paths, masks):
                    base = os.path.splitext(os.path.basename(path))[0]
                    mask_path = os.path.join(mask_output_folder, f"{base}_mask.png")
                    cv2.imwrite(mask_path, mask)
                    wm = int(np.count_nonzero(mask))
                    if wm == 0:
                        continue
                    processed.append({"image_path": path, "mask_path": mask_path,
                                      "watermark_ratio": wm / float(mask.shape[0] * mask.shape[1])})
        return processed
