"""batched watermark_filter / model_selector against the reference scripts' behaviour
(reference src/scripts/watermark_filter.py:110-196, src/scripts/model_selector.py:43-197) restated with the fp32
oracle network and cv2."""
import glob
import json
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
from tests.fixtures import synthetic_watermark_batch, train_fixture
from unet_watermark_b200.config import get_cfg_defaults
from unet_watermark_b200.scripts.model_selector import ModelSelector, calculate_watermark_metrics
from unet_watermark_b200.scripts.watermark_filter import WatermarkFilter

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(cuda_device, tmp_path_factory):
    ref, _ = train_fixture(cuda_device, steps=150, size=128, batch=16, seed=0)
    root = tmp_path_factory.mktemp("n3")
    cfg = get_cfg_defaults()
    cfg.defrost()
    cfg.MODEL.NAME = "Unet"; cfg.MODEL.ENCODER_WEIGHTS = None; cfg.DATA.IMG_SIZE = 128
    cfg.freeze()
    models = root / "models"
    models.mkdir()
    torch.save({"epoch": 7, "model_state_dict": ref.state_dict(), "val_loss": 0.2, "val_metrics": {}}, models / "a.pth")
    torch.save(O.build("resnet34", seed=3, random_bn=True).state_dict(), models / "b.pth")          # bare state dict
    (models / "broken.pth").write_bytes(b"not a checkpoint")
    imgs = root / "imgs"
    imgs.mkdir()
    _, u8, _ = synthetic_watermark_batch(6, 160, seed=11)
    clean = (O.image_like_input(2, 160, seed=5, normalise=False) * 0.5 * 255).to(torch.uint8).permute(0, 2, 3, 1).numpy()
    sizes = [(160, 160), (128, 128), (200, 120), (96, 224), (128, 128), (300, 300)]
    for i, (w, h) in enumerate(sizes):
        cv2.imwrite(str(imgs / f"wm_{i}.png"), cv2.cvtColor(cv2.resize(u8[i].numpy(), (w, h), interpolation=cv2.INTER_AREA), cv2.COLOR_RGB2BGR))
    for i in range(2):
        cv2.imwrite(str(imgs / f"clean_{i}.jpg"), cv2.cvtColor(clean[i], cv2.COLOR_RGB2BGR))
    return ref, cfg, root, imgs, models


def _ref_filter_mask(ref, path, cfg):
    """reference watermark_filter.predict_mask (:110-157) with the oracle network."""
    image = cv2.imread(path)
    rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
    x = O.val_transform(rgb, cfg.DATA.IMG_SIZE).unsqueeze(0)
    with torch.no_grad():
        prob = torch.sigmoid(ref(x))[0, 0].numpy()
    m = cv2.resize(prob, (image.shape[1], image.shape[0]))
    b = (m > cfg.PREDICT.THRESHOLD).astype(np.uint8) * 255
    if cfg.PREDICT.POST_PROCESS:
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
        b = cv2.morphologyEx(cv2.morphologyEx(b, cv2.MORPH_OPEN, k, iterations=1), cv2.MORPH_CLOSE, k, iterations=1)
    return b


def test_watermark_filter_matches_reference(setup, cuda_device, tmp_path):
    ref, cfg, root, imgs, models = setup
    f = WatermarkFilter(str(models / "a.pth"), config=cfg, device=cuda_device, watermark_threshold=0.001, batch_size=4)
    assert f.model_info["epoch"] == 7
    files = sorted(glob.glob(str(imgs / "*")))
    ratios = f.watermark_ratios(files)
    total = agree = 0
    for p, r in zip(files, ratios):
        want = _ref_filter_mask(ref.cpu(), p, cfg)
        got = f.predict_mask(p)
        total += want.size; agree += int((want == got).sum())
        assert abs(r - np.count_nonzero(got) / got.size) < 1e-12
        has, ratio = f.has_watermark(p)
        assert ratio == r and has == (r >= 0.001)
    assert agree / total >= 0.998, agree / total
    # filter_images: dry run touches nothing; a real run moves the images without watermark
    work = tmp_path / "work"
    work.mkdir()
    for p in files:
        (work / os.path.basename(p)).write_bytes(open(p, "rb").read())
    (work / "notes.txt").write_text("x")
    st = f.filter_images(str(work), str(tmp_path / "nowm"), dry_run=True)
    assert st["total"] == 8 and st["moved"] == 0 and len(list(work.iterdir())) == 9
    st2 = f.filter_images(str(work), str(tmp_path / "nowm"))
    assert st2["with_watermark"] == st["with_watermark"] and st2["moved"] == st["without_watermark"]
    assert st["with_watermark"] >= 5                                  # the trained fixture finds its watermarks
    assert len(list((tmp_path / "nowm").iterdir())) == st2["moved"]


def test_model_selector_metrics_and_report(setup, cuda_device, tmp_path):
    ref, cfg, root, imgs, models = setup
    out = tmp_path / "sel"
    sel = ModelSelector(str(imgs), str(models), str(out), num_samples=5, device=cuda_device, config=cfg)
    res = sel.run_evaluation()
    assert set(res["models"]) == {"a.pth", "b.pth", "broken.pth"} and res["num_samples"] == 5
    assert res["models"]["broken.pth"]["statistics"] is None and "load_error" in res["models"]["broken.pth"]
    a = res["models"]["a.pth"]
    assert a["statistics"]["successful_predictions"] == 5 and a["model_info"]["epoch"] == 7
    for p in a["predictions"]:
        m = cv2.imread(p["mask_path"], cv2.IMREAD_GRAYSCALE)
        n, labels, stats, _ = cv2.connectedComponentsWithStats(m.astype(np.uint8))          # reference :178
        want = {"watermark_ratio": float(np.sum(m > 0) / m.size), "watermark_pixels": int(np.sum(m > 0)),
                "total_pixels": int(m.size), "num_components": int(n - 1),
                "max_component_area": int(stats[1:, cv2.CC_STAT_AREA].max()) if n > 1 else 0,
                "max_component_ratio": float(stats[1:, cv2.CC_STAT_AREA].max() / m.size) if n > 1 else 0}
        assert p["metrics"] == want
        assert calculate_watermark_metrics(m, m.shape) == want
    assert json.load(open(out / "model_selection_results.json"))["models"]["a.pth"]["statistics"] == a["statistics"]
