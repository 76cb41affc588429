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


def _reference_step1(ref, path, img_size, threshold, sigmoid, post_process=True, mask_type="auto", ipp=True):
    """reference src/predict.py:588-634 for one file, restated with the fp32 oracle network and cv2 (the calls of
    the reference, in its order)."""
    import cv2
    from tests import cv2_reference as R
    image = cv2.imread(path)
    rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
    x = O.val_transform(rgb, img_size).unsqueeze(0)
    with torch.no_grad():
        out = ref(x)
        out = torch.sigmoid(out) if sigmoid else out
    mask = out[0, 0].numpy()
    h0, w0 = image.shape[:2]
    cv2.ipp.setUseIPP(ipp)
    try:
        resized = cv2.resize(mask, (w0, h0))
    finally:
        cv2.ipp.setUseIPP(True)
    binary = (resized > threshold).astype(np.uint8) * 255
    if not post_process:
        return binary, None
    t = R.detect_watermark_type(rgb, binary) if mask_type == "auto" else mask_type
    return R.optimize_mask(binary, t), t


def test_predict_folder_matches_reference_pipeline(trained, cuda_device, tmp_path):
    """reference step 1 (src/predict.py:588-634) restated with the oracle + cv2 vs WatermarkPredictor on files of
    different sizes: thresholded masks, detected watermark types and post-processed masks."""
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
    inp = tmp_path / "in"
    inp.mkdir()
    _, u8, _ = synthetic_watermark_batch(6, 160, seed=5)
    sizes = [(160, 160), (128, 128), (200, 120), (96, 224), (128, 128), (64, 64)]      # (w, h); 64x64 = exact 2x reduction
    for i, (w, h) in enumerate(sizes):
        img = cv2.resize(u8[i].numpy(), (w, h), interpolation=cv2.INTER_AREA)
        ext = "jpg" if i == 2 else "png"
        cv2.imwrite(str(inp / f"img_{i}.{ext}"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    (inp / "notes.txt").write_text("not an image")
    files = sorted(glob.glob(str(inp / "img_*")))

    def run(out_dir, **kw):
        pred = WatermarkPredictor(str(ckpt), config=cfg, device=cuda_device, batch_size=4, **kw)
        res = pred.step1_batch_predict_watermark_masks(str(inp), str(out_dir))
        written = sorted(glob.glob(str(out_dir / "*_mask.png")))
        assert [os.path.basename(p) for p in written] == [f"img_{i}_mask.png" for i in range(6)]
        return pred, res, [cv2.imread(p, cv2.IMREAD_GRAYSCALE) for p in written]

    # (a) raw thresholded masks, both conventions: bf16 network vs fp32 oracle
    for sig in (True, False):
        out = tmp_path / f"out_raw_{int(sig)}"
        pred, res, got = run(out, sigmoid=sig, post_process=False)
        total = agree = 0
        for i, (w, h) in enumerate(sizes):
            want, _ = _reference_step1(ref, files[i], 128, 0.5, sig, post_process=False)
            assert got[i].shape == (h, w) and set(np.unique(got[i]).tolist()) <= {0, 255}
            total += got[i].size
            agree += int((got[i] == want).sum())
        assert agree / total >= 0.998, (sig, agree / total)
    assert pred.model_info["epoch"] == 3
    # (b) the reference's step 1 as it runs: type detection + _optimize_mask
    out = tmp_path / "out_post"
    pred, res, got = run(out, sigmoid=False)
    total = agree = 0
    for i, (w, h) in enumerate(sizes):
        want, t = _reference_step1(ref, files[i], 128, 0.5, False)
        total += got[i].size
        agree += int((got[i] == want).sum())
    assert agree / total >= 0.995, agree / total
    assert all(0 < r["watermark_ratio"] <= 1 for r in res)
    # second run: everything already has a mask -> nothing to do (reference :140-144)
    assert pred.step1_batch_predict_watermark_masks(str(inp), str(out)) == []
    # single-image API (fixed type, reference :303-368) and rank sharding
    fixed = tmp_path / "out_fixed"
    predw, _, gotw = run(fixed, sigmoid=False, mask_type="watermark")
    m0 = predw.predict_mask(files[2])
    assert np.array_equal(m0, gotw[2])
    out2 = tmp_path / "out_r1"
    predw.step1_batch_predict_watermark_masks(str(inp), str(out2), rank=1, world_size=2)
    assert sorted(os.path.basename(p) for p in glob.glob(str(out2 / "*_mask.png"))) == ["img_1_mask.png", "img_3_mask.png", "img_5_mask.png"]


def test_gpu_pre_and_post_processing_is_bit_exact_given_the_same_network_output(trained, cuda_device, tmp_path):
    """Everything around the network is integer / byte work and must be BIT-exact: feed the pipeline's own network
    output through the reference's cv2 sequence (resize with OpenCV's algorithm, threshold, type detection,
    _optimize_mask) and compare with what the GPU kernels produced from the same output."""
    import cv2
    from tests import cv2_reference as R
    from unet_watermark_b200 import imgproc
    ref, _ = trained
    cfg = get_cfg_defaults()
    cfg.defrost()
    cfg.MODEL.NAME = "Unet"; cfg.MODEL.ENCODER_WEIGHTS = None; cfg.DATA.IMG_SIZE = 128
    cfg.freeze()
    ckpt = tmp_path / "m.pth"
    torch.save(ref.state_dict(), ckpt)
    pred = WatermarkPredictor(str(ckpt), config=cfg, device=cuda_device, batch_size=8)
    _, u8, _ = synthetic_watermark_batch(5, 192, seed=9)
    sizes = [(192, 192), (300, 170), (128, 128), (64, 64), (97, 411)]
    bgr = [cv2.cvtColor(cv2.resize(u8[i].numpy(), s, interpolation=cv2.INTER_AREA), cv2.COLOR_RGB2BGR) for i, s in enumerate(sizes)]
    masks, types = pred._masks_for_images(bgr, 0.5)
    # the same network output, taken from the GPU
    src = imgproc.RaggedBatch(sizes, channels=3, device=cuda_device)
    buf = torch.zeros(src.total, dtype=torch.uint8)
    for i, b in enumerate(bgr):
        d = src.host[i]
        buf[d.offset:d.offset + d.pitch * d.height] = torch.from_numpy(b.reshape(-1))
    x = imgproc.resize_bilinear_u8(buf.to(cuda_device), src, 128, 128, swap_rb=True)
    for i, b in enumerate(bgr):       # A.Resize of get_val_transform, bit for bit
        want = cv2.resize(cv2.cvtColor(b, cv2.COLOR_BGR2RGB), (128, 128), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(x[i].cpu().numpy(), want)
    maps = pred.model(x)[:, 0].cpu().numpy()
    cv2.ipp.setUseIPP(False)          # OpenCV's own float resize (the pip wheel's IPP route differs by ~1e-4)
    try:
        for i, (w, h) in enumerate(sizes):
            binary = (cv2.resize(maps[i], (w, h)) > 0.5).astype(np.uint8) * 255
            rgb = cv2.cvtColor(bgr[i], cv2.COLOR_BGR2RGB)
            t = R.detect_watermark_type(rgb, binary)
            assert t == types[i]
            assert np.array_equal(masks[i], R.optimize_mask(binary, t)), (i, sizes[i], t)
    finally:
        cv2.ipp.setUseIPP(True)
