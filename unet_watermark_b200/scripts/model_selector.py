"""Model selector: run several checkpoints over a random sample of images and compare their masks.

Mirror of reference src/scripts/model_selector.py (``evaluate_single_model`` :43-169, ``calculate_watermark_metrics``
:171-197, ``ModelSelector`` :199-582) on the batched B200 predictor: every model processes the sample in batches
(decode on CPU threads, everything else on the GPU), and the mask statistics (pixels, 8-connected components, largest
component) come from the GPU connected-components kernels instead of ``cv2.connectedComponentsWithStats``.

    python -m unet_watermark_b200.scripts.model_selector --input_dir DIR --model_dir DIR --output_dir DIR
        [--num_samples 10] [--config CFG.yaml] [--device cuda]
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import random
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime
from pathlib import Path
from typing import Dict, List

import numpy as np
import torch

from .. import imgproc
from ..predict import WatermarkPredictor

logger = logging.getLogger(__name__)


def calculate_watermark_metrics(mask, image_shape, summary=None):
    """reference :171-197.  ``summary`` = (pixels, components, largest area) from the GPU, else computed from the mask
    with the same kernels."""
    total_pixels = int(image_shape[0] * image_shape[1])
    if summary is None:
        dev = torch.device("cuda", torch.cuda.current_device())
        h, w = mask.shape[:2]
        rb = imgproc.RaggedBatch([(w, h)], channels=1, device=dev)
        buf = torch.from_numpy(np.ascontiguousarray((np.asarray(mask) > 0).astype(np.uint8) * 255).reshape(-1)).to(dev)
        summary = imgproc.mask_component_summary(buf, rb)[0]
    pixels, comps, largest = (int(v) for v in summary)
    return {"watermark_ratio": float(pixels / total_pixels), "watermark_pixels": pixels, "total_pixels": total_pixels,
            "num_components": comps, "max_component_area": largest,
            "max_component_ratio": float(largest / total_pixels) if comps > 0 else 0}


def evaluate_single_model(args):
    """reference :43-169: one checkpoint over the selected images -> (model_name, results)."""
    import cv2
    model_path, selected_images, config_path, device, output_dir = args[:5]
    cfg = args[5] if len(args) > 5 else None
    model_name = os.path.basename(model_path)
    try:
        predictor = WatermarkPredictor(model_path=model_path, config_path=config_path, config=cfg, device=device,
                                       mask_type="watermark")          # predict_mask(image_path): mask_type='watermark'
        results = {"model_path": model_path, "model_name": model_name, "model_info": predictor.model_info, "predictions": []}
        thr = float(getattr(predictor.cfg.PREDICT, "THRESHOLD", 0.5))
        bs = predictor.batch_size
        with ThreadPoolExecutor(max_workers=predictor.num_workers) as pool:
            for i0 in range(0, len(selected_images), bs):
                paths = selected_images[i0:i0 + bs]
                images = list(pool.map(predictor._decode, paths))
                ok = [(p, im) for p, im in zip(paths, images) if im is not None]
                masks, summaries = [], []
                err = None
                if ok:
                    try:
                        masks, _ = predictor._masks_for_images([im for _, im in ok], thr, pool=pool)
                        dev = predictor.device
                        rb = imgproc.RaggedBatch([(m.shape[1], m.shape[0]) for m in masks], channels=1, device=dev)
                        packed = torch.zeros(rb.total, dtype=torch.uint8)
                        for k, m in enumerate(masks):
                            rb.view(packed, k).copy_(torch.from_numpy(m))
                        summaries = imgproc.mask_component_summary(packed.to(dev), rb)
                    except Exception as e:  # noqa: BLE001
                        err = str(e)
                done = {p: (m, s) for (p, _), m, s in zip(ok, masks, summaries)} if err is None else {}
                for p, im in zip(paths, images):
                    name = os.path.basename(p)
                    if p in done:
                        m, s = done[p]
                        mask_path = os.path.join(output_dir, f"{Path(name).stem}_{model_name.replace('.pth', '')}_mask.png")
                        cv2.imwrite(mask_path, m)
                        results["predictions"].append({"image_name": name, "image_path": p, "mask_path": mask_path,
                                                       "metrics": calculate_watermark_metrics(m, im.shape[:2], s),
                                                       "success": True, "error": None})
                    else:
                        results["predictions"].append({"image_name": name, "image_path": p, "mask_path": None, "metrics": None,
                                                       "success": False, "error": err or f"无法读取图像: {p}"})
        good = [p for p in results["predictions"] if p["success"]]
        ratios = [p["metrics"]["watermark_ratio"] for p in good]
        n = len(results["predictions"])
        results["statistics"] = {
            "total_predictions": n, "successful_predictions": len(good), "failed_predictions": n - len(good),
            "avg_watermark_ratio": float(np.mean(ratios)) if ratios else 0.0,
            "std_watermark_ratio": float(np.std(ratios)) if ratios else 0.0,
            "min_watermark_ratio": float(np.min(ratios)) if ratios else 0.0,
            "max_watermark_ratio": float(np.max(ratios)) if ratios else 0.0,
            "detection_rate": float(np.mean([1 if r > 0.001 else 0 for r in ratios])) if ratios else 0.0}
        return model_name, results
    except Exception as e:  # noqa: BLE001 - reference :158-169
        logger.error("加载模型 %s 失败: %s", model_name, e)
        return model_name, {"model_path": model_path, "model_name": model_name, "model_info": None, "predictions": [],
                            "statistics": None, "load_error": str(e)}


class ModelSelector:
    """reference :199-582 (serial over models: one GPU, each model batched over the images)."""

    IMAGE_EXT = [".png", ".jpg", ".jpeg", ".bmp", ".tiff"]

    def __init__(self, input_dir, model_dir, output_dir, num_samples=10, config_path=None, device="cuda", config=None):
        self.input_dir, self.model_dir, self.output_dir = input_dir, model_dir, output_dir
        self.num_samples, self.config_path, self.device, self.config = num_samples, config_path, device, config
        os.makedirs(output_dir, exist_ok=True)
        self.image_paths = self._get_image_paths()
        self.model_paths = self._get_model_paths()
        if not self.image_paths:
            raise ValueError(f"在 {input_dir} 中未找到图片文件")
        if not self.model_paths:
            raise ValueError(f"在 {model_dir} 中未找到模型文件")

    def _get_image_paths(self):
        if os.path.isdir(self.input_dir):
            return sorted(os.path.join(self.input_dir, f) for f in os.listdir(self.input_dir)
                          if any(f.lower().endswith(e) for e in self.IMAGE_EXT))
        if os.path.isfile(self.input_dir) and any(self.input_dir.lower().endswith(e) for e in self.IMAGE_EXT):
            return [self.input_dir]
        return []

    def _get_model_paths(self):
        if os.path.isdir(self.model_dir):
            return sorted(os.path.join(r, f) for r, _, fs in os.walk(self.model_dir) for f in fs if f.endswith(".pth"))
        if os.path.isfile(self.model_dir) and self.model_dir.endswith(".pth"):
            return [self.model_dir]
        return []

    def _select_random_images(self):
        if len(self.image_paths) <= self.num_samples:
            return list(self.image_paths)
        return random.sample(self.image_paths, self.num_samples)

    def run_evaluation(self) -> Dict:
        selected = self._select_random_images()
        results = {"timestamp": datetime.now().isoformat(), "input_dir": self.input_dir, "model_dir": self.model_dir,
                   "num_samples": len(selected), "selected_images": selected, "models": {}}
        for mp in self.model_paths:
            name, res = evaluate_single_model((mp, selected, self.config_path, self.device, self.output_dir, self.config))
            results["models"][name] = res
        with open(os.path.join(self.output_dir, "model_selection_results.json"), "w", encoding="utf-8") as f:
            json.dump(results, f, indent=2, ensure_ascii=False)
        return results


def main(argv=None):
    ap = argparse.ArgumentParser(description="模型挑选脚本 (B200)")
    ap.add_argument("--input_dir", required=True)
    ap.add_argument("--model_dir", required=True)
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--num_samples", type=int, default=10)
    ap.add_argument("--config", type=str, default=None)
    ap.add_argument("--device", type=str, default="cuda")
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    dev = f"cuda:{torch.cuda.current_device()}" if a.device in ("cuda", "auto") else a.device
    res = ModelSelector(a.input_dir, a.model_dir, a.output_dir, a.num_samples, a.config, dev).run_evaluation()
    for name, r in res["models"].items():
        st = r.get("statistics")
        print(name, "load_error: " + r["load_error"] if st is None else
              f"ok {st['successful_predictions']}/{st['total_predictions']} avg ratio {st['avg_watermark_ratio']:.6f} "
              f"detection {st['detection_rate']:.2%}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
