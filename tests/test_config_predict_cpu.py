"""Config mirror, checkpoint loading, file discovery and rank sharding (host logic, no GPU)."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from unet_watermark_b200 import config as C
from unet_watermark_b200 import predict as P
from unet_watermark_b200.cli import build_parser


def test_defaults_match_reference():
    cfg = C.get_cfg_defaults()
    assert cfg.MODEL.NAME == "UnetPlusPlus" and cfg.MODEL.ENCODER_NAME == "resnet34"
    assert cfg.MODEL.DECODER_CHANNELS == [256, 128, 64, 32, 16] and cfg.MODEL.ACTIVATION is None
    assert cfg.DATA.IMG_SIZE == 512 and cfg.PREDICT.THRESHOLD == 0.5 and cfg.DEVICE == "cpu"
    assert cfg.LOSS.BCE_WEIGHT == 0.5 and cfg.LOSS.DICE_WEIGHT == 0.5 and cfg.TRAIN.LR == 1e-4


def test_merge_freeze_clone(tmp_path):
    y = tmp_path / "c.yaml"
    y.write_text("MODEL:\n  NAME: Unet\n  ENCODER_NAME: resnet50\nDATA:\n  IMG_SIZE: 768\n")
    cfg = C.get_cfg_defaults()
    C.update_config(cfg, str(y))
    assert cfg.MODEL.NAME == "Unet" and cfg.MODEL.ENCODER_NAME == "resnet50" and cfg.DATA.IMG_SIZE == 768
    with pytest.raises(AttributeError):
        cfg.MODEL.NAME = "x"
    c2 = cfg.clone()
    c2.defrost()
    c2.MODEL.NAME = "FPN"
    assert cfg.MODEL.NAME == "Unet" and C.get_cfg_defaults().MODEL.NAME == "UnetPlusPlus"
    bad = tmp_path / "bad.yaml"
    bad.write_text("MODEL:\n  NOPE: 1\n")
    with pytest.raises(KeyError):
        C.update_config(C.get_cfg_defaults(), str(bad))
    c2.merge_from_list(["PREDICT.THRESHOLD", "0.3", "MODEL.ACTIVATION", "sigmoid"])
    assert c2.PREDICT.THRESHOLD == 0.3 and c2.MODEL.ACTIVATION == "sigmoid"
    assert hasattr(cfg.MODEL, "ENCODER_DEPTH") and not hasattr(cfg.MODEL, "MISSING")


def test_checkpoint_formats_and_yacs_shim(tmp_path):
    """reference trainer checkpoints pickle a yacs CfgNode under 'config' (src/train.py:428-435)."""
    C.install_yacs_shim()
    from yacs.config import CfgNode
    cfg = CfgNode({"MODEL": {"NAME": "Unet"}})
    sd = {"segmentation_head.0.bias": torch.ones(1)}
    p1 = tmp_path / "wrapped.pth"
    torch.save({"epoch": 7, "model_state_dict": sd, "val_loss": 0.25, "val_metrics": {"iou": 0.9}, "config": cfg}, p1)
    state, info = P.load_checkpoint_state(str(p1))
    assert info["epoch"] == 7 and info["val_metrics"]["iou"] == 0.9 and torch.equal(state["segmentation_head.0.bias"], sd["segmentation_head.0.bias"])
    p2 = tmp_path / "bare.pth"
    torch.save(sd, p2)
    state, info = P.load_checkpoint_state(str(p2))
    assert info["epoch"] == "Unknown" and "segmentation_head.0.bias" in state
    with pytest.raises(FileNotFoundError):
        P.load_checkpoint_state(str(tmp_path / "missing.pth"))
    assert pickle.loads(pickle.dumps(cfg)).MODEL.NAME == "Unet"


def test_predictor_requires_cuda(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.WatermarkPredictor(str(tmp_path / "m.pth"), device="cpu")


def test_get_image_files_rules(tmp_path):
    inp, out = tmp_path / "in", tmp_path / "out"
    inp.mkdir(); out.mkdir()
    for n in ("b.jpg", "a.PNG", "c.webp", "d.txt", "e.jpeg", "f.bmp", "g.tiff"):
        (inp / n).write_bytes(b"x")
    (out / "b_mask.png").write_bytes(b"x")                       # already processed -> skipped
    pred = object.__new__(P.WatermarkPredictor)
    files = pred._get_image_files(str(inp), str(out))
    assert [os.path.basename(f) for f in files] == ["a.PNG", "c.webp", "e.jpeg", "f.bmp", "g.tiff"]
    assert len(pred._get_image_files(str(inp), None)) == 6
    assert len(pred._get_image_files(str(inp), str(out), limit=2)) == 2


def test_rank_partition_is_stable_when_other_ranks_already_wrote_masks(tmp_path):
    """shard the full sorted list first, filter inside the shard; --limit picks the same subset on all ranks."""
    inp, out = tmp_path / "in", tmp_path / "out"
    inp.mkdir(); out.mkdir()
    names = [f"img_{i:02d}.png" for i in range(10)]
    for n in names:
        (inp / n).write_bytes(b"x")
    pred = object.__new__(P.WatermarkPredictor)
    r0 = pred._get_image_files(str(inp), str(out), rank=0, world_size=2)
    for f in r0[:3]:                                              # rank 0 is ahead: three of its masks exist already
        (out / (os.path.splitext(os.path.basename(f))[0] + "_mask.png")).write_bytes(b"x")
    r1 = pred._get_image_files(str(inp), str(out), rank=1, world_size=2)       # rank 1 lists the folder late
    assert [os.path.basename(f) for f in r0] == names[0::2]
    assert [os.path.basename(f) for f in r1] == names[1::2]      # unchanged by rank 0's progress
    assert [os.path.basename(f) for f in pred._get_image_files(str(inp), str(out), rank=0, world_size=2)] == names[6::2]
    a = pred._get_image_files(str(inp), None, limit=4, rank=0, world_size=2)
    b = pred._get_image_files(str(inp), None, limit=4, rank=1, world_size=2)
    assert len(a) + len(b) == 4 and not set(a) & set(b)


def test_shard_for_rank_partitions():
    items = list(range(11))
    shards = [P.shard_for_rank(items, r, 4) for r in range(4)]
    assert sorted(sum(shards, [])) == items
    assert max(map(len, shards)) - min(map(len, shards)) <= 1
    assert P.shard_for_rank([], 0, 2) == []
    with pytest.raises(ValueError):
        P.shard_for_rank(items, 4, 4)


def _gloo_worker(rank, world, port, n_items, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.shard_for_rank([f"img_{i:03d}.png" for i in range(n_items)], rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # test-only collective: the data path itself has none
    # per-rank image counts -> whole-job total, as bench.py aggregates throughput
    t = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        q.put((gathered, int(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_covers_every_image_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 13, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 13
    assert sorted(gathered[0] + gathered[1]) == [f"img_{i:03d}.png" for i in range(13)]
    assert not set(gathered[0]) & set(gathered[1])


def test_cli_surface():
    p = build_parser()
    a = p.parse_args(["predict", "--input", "x", "--output", "y", "--model", "m.pth", "--save-mask", "--limit", "3"])
    assert (a.command, a.input, a.output, a.model, a.limit, a.sigmoid) == ("predict", "x", "y", "m.pth", 3, False)
    a = p.parse_args(["predict", "--sigmoid", "--threshold", "0.3", "--batch-size", "4"])
    assert a.sigmoid and a.threshold == 0.3 and a.batch_size == 4
    assert p.parse_args(["predict", "--no-sigmoid"]).sigmoid is False      # compat flag, default convention


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn as nn
    from unet_watermark_b200.training import GradBuckets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                           # identical replicas
    net = nn.Sequential(nn.Linear(7, 5), nn.ReLU(), nn.Linear(5, 3), nn.ReLU(), nn.Linear(3, 1))
    import copy
    plain = copy.deepcopy(net)                                     # same weights, no buckets: this rank's own gradient
    buckets = GradBuckets(net.parameters(), bucket_mb=1e-4)        # ~26 floats per bucket: several buckets
    g = torch.Generator().manual_seed(100 + rank)                  # different data per rank
    x, y = torch.randn(6, 7, generator=g), torch.randn(6, 1, generator=g)
    ((plain(x) - y) ** 2).mean().backward()
    # (read from the un-bucketed copy: the bucketed net's all-reduces start inside backward() and run in the background,
    # so its p.grad may already hold partial sums when backward() returns)
    local = [p.grad.flatten().tolist() for p in plain.parameters()]    # plain lists: they cross process boundaries
    buckets.zero_grad()
    ((net(x) - y) ** 2).mean().backward()
    buckets.finish()
    avg = [p.grad.flatten().tolist() for p in net.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        q.put((len(buckets.buckets), avg, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_bucketed_gradient_allreduce_averages():
    """The exchange step of the optional training path (config 5): every rank ends with the mean gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_buckets, avg, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_buckets >= 2
    for i, a in enumerate(avg):
        want = (torch.tensor(gathered[0][i]) + torch.tensor(gathered[1][i])) / 2
        assert torch.allclose(torch.tensor(a), want, atol=1e-6)
    assert any(gathered[0][i] != gathered[1][i] for i in range(len(avg)))      # the ranks really saw different data


def test_device_prefetcher_passes_batches_through_on_cpu():
    from unet_watermark_b200.training import DevicePrefetcher
    batches = [(torch.full((2, 3, 4, 4), float(i)), torch.full((2, 4, 4), i, dtype=torch.long)) for i in range(5)]
    got = [(x.clone(), t.clone()) for x, t in DevicePrefetcher(iter(batches), "cpu")]
    assert len(got) == 5
    for (x, t), (gx, gt) in zip(batches, got):
        assert torch.equal(x, gx) and torch.equal(t, gt)
    assert list(DevicePrefetcher(iter(()), "cpu")) == []
