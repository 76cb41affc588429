"""Trained-fixture mask agreement (>= 99.9 %) and the file-level predict surface, on the GPU."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
from tests.fixtures import synthetic_watermark_batch, train_fixture
from unet_watermark_b200.config import CfgNode, get_cfg_defaults, install_yacs_shim
from unet_watermark_b200.predict import WatermarkPredictor
from unet_watermark_b200.unet_model import Unet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained(cuda_device):
    model, loss = train_fixture(cuda_device, steps=200, size=128, batch=16, seed=0)
    return model, loss


def test_trained_fixture_mask_agreement_999(trained, cuda_device):
    """North-star criterion: binarised masks agree with the reference on >= 99.9 % of pixels."""
    ref, loss = trained
    m = Unet("resnet34", encoder_weights=None)
    m.load_state_dict(ref.state_dict(), strict=True)
    m = m.to(cuda_device).eval()
    x, u8, target = synthetic_watermark_batch(8, 256, seed=77)
    with torch.no_grad():
        y32 = ref(x)
    ref_mask = O.binarize(y32[:, 0], 0.5, sigmoid=True)
    # the fixture must be meaningful: it segments, and both classes are present
    iou = ((ref_mask > 0) & (target[:, 0] > 0)).sum().item() / max(((ref_mask > 0) | (target[:, 0] > 0)).sum().item(), 1)
    assert iou > 0.5, f"fixture did not train (loss {loss}, IoU {iou})"
    assert 0.01 < (ref_mask > 0).float().mean().item() < 0.6
    mask, logits = m.predict_mask(x.to(cuda_device), 0.5, sigmoid=True, return_logits=True)
    agree = (mask.cpu() == ref_mask).float().mean().item()
    d = (logits.cpu() - y32).abs()
    print(f"trained fixture: IoU {iou:.3f}, mask agreement {agree:.6f}, logits max|d| {d.max().item():.4f} "
          f"(absmax {y32.abs().max().item():.2f}), mean|d| {d.mean().item():.5f} (std {y32.std().item():.3f})")
    assert agree >= 0.999
    assert d.max() <= 0.08 * y32.abs().max() and d.mean() <= 0.05 * y32.std()
    # same through the fused uint8 input path
    mask8 = m.predict_mask(u8.to(cuda_device), 0.5)
    assert (mask8.cpu() == ref_mask).float().mean().item() >= 0.999


def test_predict_folder_matches_reference_pipeline(trained, cuda_device, tmp_path):
    """reference step 1 (src/predict.py:588-634) restated with the oracle vs WatermarkPredictor on files."""
    import cv2
    ref, _ = trained
    install_yacs_shim()
    cfg = get_cfg_defaults()
    cfg.defrost()
    cfg.MODEL.NAME = "Unet"
    cfg.MODEL.ENCODER_WEIGHTS = None
    cfg.DATA.IMG_SIZE = 128
    cfg.freeze()
    ckpt = tmp_path / "best.pth"          # the reference trainer's dict format, CfgNode pickled inside
    torch.save({"epoch": 3, "model_state_dict": ref.state_dict(), "val_loss": 0.1, "val_metrics": {"iou": 0.8},
                "config": CfgNode(cfg.to_dict())}, ckpt)
    inp, out = tmp_path / "in", tmp_path / "out"
    inp.mkdir()
    _, u8, _ = synthetic_watermark_batch(5, 160, seed=5)
    sizes = [(160, 160), (128, 128), (200, 120), (96, 224), (128, 128)]      # (w, h)
    for i, (w, h) in enumerate(sizes):
        img = cv2.resize(u8[i].numpy(), (w, h), interpolation=cv2.INTER_AREA)
        cv2.imwrite(str(inp / f"img_{i}.png"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    (inp / "notes.txt").write_text("not an image")
    pred = WatermarkPredictor(str(ckpt), config=cfg, device=cuda_device, batch_size=2, sigmoid=True)
    assert pred.model_info["epoch"] == 3
    res = pred.step1_batch_predict_watermark_masks(str(inp), str(out))
    written = sorted(glob.glob(str(out / "*_mask.png")))
    assert [os.path.basename(p) for p in written] == [f"img_{i}_mask.png" for i in range(5)]
    total = agree = 0
    for i, (w, h) in enumerate(sizes):
        bgr = cv2.imread(str(inp / f"img_{i}.png"))
        rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        x = O.val_transform(rgb, 128).unsqueeze(0)
        with torch.no_grad():
            prob = torch.sigmoid(ref(x))[0, 0].numpy()
        want = O.resize_and_binarize(prob, (w, h), 0.5)
        got = cv2.imread(written[i], cv2.IMREAD_GRAYSCALE)
        assert got.shape == (h, w) and set(np.unique(got).tolist()) <= {0, 255}
        total += got.size
        agree += int((got == want).sum())
    assert agree / total >= 0.998, agree / total
    assert all(r["watermark_ratio"] > 0 for r in res)
    # second run: everything already has a mask -> nothing to do (reference :140-144)
    assert pred.step1_batch_predict_watermark_masks(str(inp), str(out)) == []
    # single-image API and rank sharding
    m0 = pred.predict_mask(str(inp / "img_2.png"))
    assert np.array_equal(m0, cv2.imread(written[2], cv2.IMREAD_GRAYSCALE))
    out2 = tmp_path / "out_r1"
    pred.step1_batch_predict_watermark_masks(str(inp), str(out2), rank=1, world_size=2)
    assert sorted(os.path.basename(p) for p in glob.glob(str(out2 / "*_mask.png"))) == ["img_1_mask.png", "img_3_mask.png"]
