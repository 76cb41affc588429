"""Watermark filter: keep images that contain a watermark, move / delete the others.

Mirror of reference src/scripts/watermark_filter.py (``WatermarkFilter``: :34-108 constructor / load, :110-157
``predict_mask``, :159-166 ``_post_process_mask``, :168-196 ``has_watermark``, :198-280 ``filter_images``, :282-340
CLI) with the per-image loop replaced by the batched B200 path: decode on CPU threads, resize / network / sigmoid /
bilinear upscale / threshold / 3x3 open + close on the GPU (``WatermarkPredictor._masks_for_images``).

    python -m unet_watermark_b200.scripts.watermark_filter --input_dir DIR --model_path M.pth [--config_path C.yaml]
        [--threshold 0.0001] [--no_watermark_dir DIR] [--dry_run] [--batch_size 16]
"""
from __future__ import annotations

import argparse
import logging
import os
import shutil
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import imgproc
from ..config import get_cfg_defaults, update_config
from ..predict import WatermarkPredictor

logger = logging.getLogger(__name__)

# reference :159-166: cv2.morphologyEx(OPEN) then (CLOSE) with a 3x3 ellipse, one iteration each
_POST = [(imgproc.OP_OPEN, imgproc.SHAPE_ELLIPSE, (3, 3), 1), (imgproc.OP_CLOSE, imgproc.SHAPE_ELLIPSE, (3, 3), 1)]
IMAGE_EXTENSIONS = [".jpg", ".jpeg", ".png", ".bmp", ".tiff", ".tif"]


class WatermarkFilter:
    def __init__(self, model_path, config_path=None, device="auto", watermark_threshold=0.001, batch_size: int = 16,
                 num_workers: int = 8, config=None):
        if device == "auto":
            device = "cuda" if torch.cuda.is_available() else "cpu"        # the predictor refuses anything but CUDA
        if device == "cuda":
            device = f"cuda:{torch.cuda.current_device()}"
        self.device = torch.device(device)
        self.watermark_threshold = watermark_threshold
        if config is not None:
            self.cfg = config
        else:
            self.cfg = get_cfg_defaults()
            if config_path and os.path.exists(config_path):
                update_config(self.cfg, config_path, strict=False)
        self.batch_size, self.num_workers = batch_size, num_workers
        # sigmoid convention (reference :136), the filter's own post-processing instead of _optimize_mask
        self.predictor = WatermarkPredictor(model_path, config=self.cfg, device=self.device, batch_size=batch_size,
                                            sigmoid=True, num_workers=num_workers, post_process=False)
        self.model, self.model_info = self.predictor.model, self.predictor.model_info
        logger.info("水印过滤器初始化完成 (设备 %s, 水印阈值 %s, 模型 %s)", self.device, watermark_threshold,
                    os.path.basename(str(model_path)))

    def _masks(self, images_bgr, pool=None, slot=0) -> List[np.ndarray]:
        post = _POST if self.cfg.PREDICT.POST_PROCESS else []
        masks, _ = self.predictor._masks_for_images(images_bgr, float(self.cfg.PREDICT.THRESHOLD), slot=slot, pool=pool,
                                                    morphology=post, sigmoid=True)
        return masks

    def predict_mask(self, image_path) -> np.ndarray:
        """reference :110-157."""
        image = self.predictor._decode(image_path)
        if image is None:
            raise ValueError(f"无法读取图像: {image_path}")
        return self._masks([image])[0]

    def has_watermark(self, image_path) -> Tuple[bool, float]:
        """reference :168-196."""
        try:
            mask = self.predict_mask(image_path)
            ratio = float(np.sum(mask > 0)) / float(mask.shape[0] * mask.shape[1])
            return ratio >= self.watermark_threshold, ratio
        except Exception as e:  # noqa: BLE001 - reference behaviour
            logger.error("检测图像 %s 时出错: %s", image_path, e)
            return False, 0.0

    def watermark_ratios(self, image_paths: List[str]) -> List[Optional[float]]:
        """Batched ``has_watermark``: watermark area ratio per image (None where the image cannot be read)."""
        out: List[Optional[float]] = [None] * len(image_paths)
        bs = self.batch_size
        with ThreadPoolExecutor(max_workers=self.num_workers) as pool:
            chunks = [list(range(i, min(i + bs, len(image_paths)))) for i in range(0, len(image_paths), bs)]
            pending = [pool.submit(self.predictor._decode, image_paths[i]) for i in chunks[0]] if chunks else []
            for ci, idx in enumerate(chunks):
                images = [f.result() for f in pending]
                if ci + 1 < len(chunks):
                    pending = [pool.submit(self.predictor._decode, image_paths[i]) for i in chunks[ci + 1]]
                ok = [(i, im) for i, im in zip(idx, images) if im is not None]
                if not ok:
                    continue
                try:
                    masks = self._masks([im for _, im in ok], pool=pool, slot=ci & 1)
                except Exception as e:  # noqa: BLE001
                    logger.error("检测批次时出错: %s", e)
                    continue
                for (i, _), m in zip(ok, masks):
                    out[i] = float(np.count_nonzero(m)) / float(m.shape[0] * m.shape[1])
        return out

    def filter_images(self, input_dir, no_watermark_dir=None, dry_run=False):
        """reference :198-280, batched."""
        files = []
        for ext in IMAGE_EXTENSIONS:
            files.extend(Path(input_dir).glob(f"*{ext}"))
            files.extend(Path(input_dir).glob(f"*{ext.upper()}"))
        files = sorted(set(files))
        if not files:
            logger.warning("在 %s 中未找到图像文件", input_dir)
            return {"total": 0, "with_watermark": 0, "without_watermark": 0, "moved": 0, "errors": 0}
        if no_watermark_dir and not dry_run:
            os.makedirs(no_watermark_dir, exist_ok=True)
        stats = {"total": len(files), "with_watermark": 0, "without_watermark": 0, "moved": 0, "errors": 0}
        ratios = self.watermark_ratios([str(p) for p in files])
        for path, ratio in zip(files, ratios):
            if ratio is None:                       # unreadable: the reference logs the error and counts 'no watermark' (:190-196)
                logger.error("检测图像 %s 时出错: 无法读取图像", path)
                ratio = 0.0
            try:
                if ratio >= self.watermark_threshold:
                    stats["with_watermark"] += 1
                    continue
                stats["without_watermark"] += 1
                if dry_run:
                    continue
                if no_watermark_dir:
                    shutil.move(str(path), os.path.join(no_watermark_dir, path.name))
                else:
                    os.remove(str(path))
                stats["moved"] += 1
            except Exception as e:  # noqa: BLE001
                stats["errors"] += 1
                logger.error("处理 %s 时出错: %s", path.name, e)
        return stats


def main(argv=None):
    ap = argparse.ArgumentParser(description="水印检测过滤脚本 (B200)")
    ap.add_argument("--input_dir", type=str, default="data/train/watermarked")
    ap.add_argument("--model_path", type=str, default="models/unet_watermark.pth")
    ap.add_argument("--config_path", type=str)
    ap.add_argument("--device", type=str, default="auto")
    ap.add_argument("--threshold", type=float, default=0.0001)
    ap.add_argument("--no_watermark_dir", type=str)
    ap.add_argument("--dry_run", action="store_true")
    ap.add_argument("--batch_size", type=int, default=16)
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    for what, p in (("输入目录", a.input_dir), ("模型文件", a.model_path)):
        if not os.path.exists(p):
            logger.error("%s不存在: %s", what, p)
            return 2
    if a.config_path and not os.path.exists(a.config_path):
        logger.error("配置文件不存在: %s", a.config_path)
        return 2
    f = WatermarkFilter(a.model_path, a.config_path, a.device, a.threshold, batch_size=a.batch_size)
    stats = f.filter_images(a.input_dir, a.no_watermark_dir, a.dry_run)
    logger.info("处理完成: 总图片数 %d, 有水印 %d, 无水印 %d, 移动/删除 %d, 错误 %d", stats["total"], stats["with_watermark"],
                stats["without_watermark"], stats["moved"], stats["errors"])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
