"""Command-line surface of the path (mirror of reference src/cli.py for mask inference).

The reference documents ``main.py predict --input <dir|file> --output <dir> --model <pth>``
(reference src/cli.py:352-354, README.md:106-108) but only registers ``train | repair | auto-train``
(reference src/cli.py:371,398,468).  This module implements the documented ``predict`` command with the
flags of the reference's ``repair`` step 1 (reference src/cli.py:400-409,424): ``--input --output --model
--config --device --limit`` plus ``--batch-size --threshold --sigmoid --no-post-process --mask-type``.  ``train``, ``repair`` and
``auto-train`` orchestrate external tools and are out of scope.

Multi-GPU: launch under ``torchrun --nproc-per-node N``; each rank takes every N-th file of the
sorted list (no collective on the inference path).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import torch

from .config import get_cfg_defaults, update_config
from .predict import WatermarkPredictor


def setup_device(device_str):
    """reference src/cli.py:23-43 ('auto' -> cuda if available)."""
    if device_str == "auto":
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    else:
        device = torch.device(device_str)
    print(f"使用设备: {device}")
    if device.type == "cuda":
        print(f"GPU名称: {torch.cuda.get_device_name(device)}")
        print(f"GPU内存: {torch.cuda.get_device_properties(device).total_memory / 1024**3:.1f} GB")
    return device


def predict_command(args):
    if not os.path.exists(args.input):
        print(f"错误: 输入路径不存在: {args.input}")
        return 2
    if not os.path.exists(args.model):
        print(f"错误: 模型文件不存在: {args.model}")
        return 2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev_str = args.device
    if dev_str in ("auto", "cuda") and torch.cuda.is_available():
        dev_str = f"cuda:{local_rank if world > 1 else torch.cuda.current_device()}"      # always an indexed device
    device = setup_device(dev_str)
    if device.type == "cuda":
        torch.cuda.set_device(device)

    cfg = get_cfg_defaults()
    if args.config and os.path.exists(args.config):
        update_config(cfg, args.config, strict=False)      # every YAML the reference ships loads (unknown keys kept + logged)
    elif args.config:
        print(f"警告: 配置文件不存在: {args.config}，使用默认配置")
    cfg.defrost()
    cfg.MODEL.NAME = args.model_name or cfg.MODEL.NAME
    if cfg.MODEL.NAME not in ("Unet", "UnetPlusPlus"):
        # No silent remapping of the architecture: a checkpoint only loads into the decoder it was trained with.
        print(f"错误: MODEL.NAME={cfg.MODEL.NAME!r} 不受支持: the B200 mask path implements 'Unet' (static plan, CUDA "
              f"graph) and 'UnetPlusPlus' (the reference default; operator sequence), resnet34/resnet50 encoders. "
              f"Pass --model-name to override the config.")
        return 2
    if args.encoder:
        cfg.MODEL.ENCODER_NAME = args.encoder
    if args.img_size:
        cfg.DATA.IMG_SIZE = args.img_size
    if args.threshold is not None:
        cfg.PREDICT.THRESHOLD = args.threshold
    cfg.freeze()

    os.makedirs(args.output, exist_ok=True)
    predictor = WatermarkPredictor(model_path=args.model, config=cfg, device=device, batch_size=args.batch_size,
                                   sigmoid=args.sigmoid, num_workers=args.workers,
                                   post_process=False if args.no_post_process else None, mask_type=args.mask_type)
    t0 = time.time()
    tmp = None
    if os.path.isfile(args.input):           # single image: same code path over a one-file folder view
        tmp = tempfile.mkdtemp(prefix="uwm_single_")
        os.symlink(os.path.abspath(args.input), os.path.join(tmp, os.path.basename(args.input)))
        folder = tmp
    else:
        folder = args.input
    try:
        results = predictor.step1_batch_predict_watermark_masks(folder, args.output, limit=args.limit, rank=rank,
                                                                world_size=world)
    finally:
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)
    dt = time.time() - t0
    summary = {"status": "success", "rank": rank, "world_size": world, "masks_with_watermark": len(results),
               "total_time": dt, "results": results}
    with open(os.path.join(args.output, f"predict_summary_rank{rank}.json"), "w", encoding="utf-8") as f:
        json.dump(summary, f, indent=2, ensure_ascii=False)
    print(f"预测完成: rank {rank}/{world}, 检测到水印的图片 {len(results)} 张, 用时 {dt:.2f}s")
    return 0


def build_parser():
    parser = argparse.ArgumentParser(description="水印分割系统 - B200 native UNet mask inference",
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = parser.add_subparsers(dest="command", help="可用命令")
    p = sub.add_parser("predict", help="预测水印掩码")
    p.add_argument("--input", type=str, default="data/test", help="输入图像路径或目录")
    p.add_argument("--output", type=str, default="data/result", help="输出目录")
    p.add_argument("--model", type=str, default="models/unet_watermark.pth", help="模型文件路径")
    p.add_argument("--config", type=str, default=None, help="配置文件路径 (YAML)")
    p.add_argument("--device", type=str, default="auto", help="计算设备 (默认: auto)")
    p.add_argument("--limit", type=int, default=None, help="限制处理的图片数量")
    p.add_argument("--save-mask", action="store_true", help="(compat) masks are always saved")
    p.add_argument("--batch-size", type=int, default=16)
    p.add_argument("--threshold", type=float, default=None, help="二值化阈值 (默认: cfg.PREDICT.THRESHOLD)")
    p.add_argument("--sigmoid", action="store_true",
                   help="threshold sigmoid(output) (watermark_filter.py:136-150 convention) instead of the reference "
                        "predictor's raw output > threshold (predict.py:624-625, the default)")
    p.add_argument("--no-sigmoid", action="store_true", help="(compat; the raw-output convention is the default)")
    p.add_argument("--no-post-process", action="store_true",
                   help="write the thresholded masks without the reference's _optimize_mask step (predict.py:161-301)")
    p.add_argument("--mask-type", choices=["auto", "watermark", "text", "mixed"], default="auto",
                   help="auto: detect the watermark type per image like the reference's step 1 (predict.py:414-441); "
                        "a fixed type skips the detection (as predict_mask does)")
    p.add_argument("--model-name", type=str, default=None, help="override cfg.MODEL.NAME (Unet | UnetPlusPlus; config default: UnetPlusPlus)")
    p.add_argument("--encoder", type=str, default=None, help="override cfg.MODEL.ENCODER_NAME")
    p.add_argument("--img-size", type=int, default=None, help="override cfg.DATA.IMG_SIZE")
    p.add_argument("--workers", type=int, default=8, help="CPU decode threads")
    for name in ("train", "repair", "auto-train"):
        sub.add_parser(name, help="(reference command; out of scope for the B200 hot path)")
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.command == "predict":
        return predict_command(args)
    if args.command in ("train", "repair", "auto-train"):
        print(f"'{args.command}' orchestrates components outside the B200 mask-inference path "
              "(IOPaint/OCR/training loop); use the reference for it. See DESIGN.md.")
        return 2
    parser.print_help()
    return 0


if __name__ == "__main__":
    sys.exit(main())
