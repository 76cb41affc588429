"""BASELINE config 5 (optional training step): our train-mode forward (tcgen05 convs + train-mode BatchNorm) and the
torch-autograd backward against the fp32 oracle trained the same way (reference src/train.py:82-107 semantics:
model.train(), Dice+BCE composition, Adam)."""
import copy

import pytest
import torch

from oracle import unet_oracle as O
from tests.fixtures import synthetic_watermark_batch
from unet_watermark_b200 import _lib
from unet_watermark_b200.losses import CombinedLoss, DiceLoss, dice_bce_from_config
from unet_watermark_b200.config import get_cfg_defaults
from unet_watermark_b200.training import TrainStep
from unet_watermark_b200.unet_model import Unet

pytestmark = pytest.mark.gpu


def _pair(dev, seed=0):
    ref = O.build("resnet34", seed=seed, random_bn=True).to(dev).train()
    m = Unet("resnet34", encoder_weights=None)
    m.load_state_dict(ref.state_dict(), strict=True)
    return ref, m.to(dev).train()


def test_losses_equal_oracle_restatement(cuda_device):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 1, 32, 32, generator=g).to(cuda_device)
    target = (torch.rand(3, 1, 32, 32, generator=g) > 0.7).long().to(cuda_device)
    cfg = get_cfg_defaults()
    assert abs(float(dice_bce_from_config(cfg)(logits, target)) - float(O.dice_bce_loss(logits, target.float()))) < 1e-6
    assert abs(float(DiceLoss(smooth=1e-5)(logits, target)) - float(O.dice_loss_binary(logits, target.float()))) < 1e-6
    assert float(DiceLoss(smooth=1e-5)(logits, torch.zeros_like(target))) == 0.0
    assert isinstance(dice_bce_from_config(cfg), CombinedLoss)


def test_train_mode_forward_backward_match_oracle(cuda_device):
    ref, m = _pair(cuda_device)
    x, _, t = synthetic_watermark_batch(4, 128, seed=3, device=cuda_device)
    before = _lib.load().uwm_kernel_launch_count()
    y = m(x)
    assert _lib.load().uwm_kernel_launch_count() - before >= 40          # the convs ran on the tcgen05 kernel
    assert y.requires_grad and y.shape == (4, 1, 128, 128) and y.dtype == torch.float32
    ref16 = copy.deepcopy(ref).to(torch.bfloat16).to(memory_format=torch.channels_last)      # yardstick: stock torch bf16
    y32 = ref(x)
    y16g = ref16(x.to(torch.bfloat16)).float()
    y16 = y16g.detach()
    d = (y.detach() - y32.detach()).abs()
    d16 = (y16 - y32.detach()).abs()
    # train-mode BatchNorm re-centres every layer of a random net, which amplifies bf16 rounding (SURVEY.md App. D:
    # stock bf16 is 18-31 % of the scale away from fp32 on such fixtures): the gate is the stock-bf16 distance
    assert d.mean() <= 1.5 * d16.mean() + 1e-3, (d.mean().item(), d16.mean().item())
    assert d.max() <= 2.0 * d16.max() + 1e-2, (d.max().item(), d16.max().item())
    assert d.max() <= 0.35 * y32.abs().max() and d.mean() <= 0.25 * y32.std(), (d.max().item(), d.mean().item())
    loss, loss32 = O.dice_bce_loss(y, t), O.dice_bce_loss(y32, t)
    assert abs(loss.item() - loss32.item()) <= 0.05 * abs(loss32.item()) + 1e-3
    loss.backward()
    loss32.backward()
    O.dice_bce_loss(y16g, t).backward()
    # gradients: direction against the fp32 oracle, with stock torch bf16 (same net, same data) as the yardstick -
    # through 47 train-mode-BatchNorm layers of a random net bf16 gradients of the early layers are noisy for anyone
    cos = torch.nn.functional.cosine_similarity
    ours, stock = [], []
    for (k, p), (_, q), (_, q16) in zip(m.named_parameters(), ref.named_parameters(), ref16.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, k
        if q.grad.norm() > 0 and p.numel() > 10000:
            c = cos(p.grad.flatten().float(), q.grad.flatten(), dim=0).item()
            c16 = cos(q16.grad.flatten().float(), q.grad.flatten(), dim=0).item()
            ours.append(c); stock.append(c16)
            assert c >= c16 - 0.15, (k, c, c16)
    assert sum(ours) / len(ours) >= sum(stock) / len(stock) - 0.05, (sum(ours) / len(ours), sum(stock) / len(stock))
    assert sum(ours) / len(ours) > 0.7
    # the layers next to the loss see little accumulated noise
    last = dict(m.named_parameters())["decoder.blocks.4.conv1.0.weight"].grad.flatten().float()
    assert cos(last, dict(ref.named_parameters())["decoder.blocks.4.conv1.0.weight"].grad.flatten(), dim=0).item() > 0.97
    # train-mode BatchNorm updated the running statistics like the oracle's
    for (k, b), (_, b32) in zip(m.named_buffers(), ref.named_buffers()):
        if k.endswith("running_mean"):
            assert (b - b32).abs().max() <= 0.05 * b32.abs().max() + 0.02, k
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(b32) == 1


def test_train_steps_reduce_the_loss_and_eval_path_sees_new_weights(cuda_device):
    torch.manual_seed(0)
    m = Unet("resnet34", encoder_weights=None).to(cuda_device)
    ts = TrainStep(m, lr=1e-3)
    losses = []
    for it in range(40):
        x, _, t = synthetic_watermark_batch(8, 128, seed=500 + it, device=cuda_device)
        losses.append(float(ts.step(x, t[:, 0].long())))
    assert sum(losses[-5:]) / 5 < 0.8 * sum(losses[:5]) / 5, (losses[:5], losses[-5:])
    # the inference plan re-folds BatchNorm from the trained parameters (weight generation tracking)
    m.eval()
    x, u8, t = synthetic_watermark_batch(4, 128, seed=999, device=cuda_device)
    y_eval = m(x)
    ref = O.Unet("resnet34").to(cuda_device)
    ref.load_state_dict(m.state_dict())
    ref.eval()
    with torch.no_grad():
        y32 = ref(x)
    d = (y_eval - y32).abs()
    assert d.max() <= 0.08 * y32.abs().max() + 0.05 and d.mean() <= 0.05 * y32.std() + 0.01
    snapshot = copy.deepcopy(m.state_dict())
    ts.step(x, t[:, 0].long())
    m.eval()
    assert not torch.equal(m(x), y_eval)                       # one more step changed the weights -> new plan weights
    assert any(not torch.equal(v, snapshot[k]) for k, v in m.state_dict().items() if "num_batches" not in k)


# every data-gradient geometry of the Unet-resnet34 training step that runs on the tcgen05 kernel (gy channels = Cout of
# the forward conv, gx channels = its Cin; the decoder conv1 gradients fan OUT to the concat's 768 / 384 / 192 / 128 channels)
# (name, n, h, w, cin, cout, k)
DGRAD_CASES = [
    ("l1_64_64",        2, 32, 32,  64,  64, 3),
    ("l2_128_128",      2, 16, 16, 128, 128, 3),
    ("l3_256_256",      2,  8,  8, 256, 256, 3),
    ("l4_512_512",      4,  4,  4, 512, 512, 3),
    ("d0_conv1_768_256", 2, 8,  8, 768, 256, 3),
    ("d1_conv1_384_128", 1, 16, 16, 384, 128, 3),
    ("d2_conv1_192_64", 1, 32, 32, 192,  64, 3),
    ("d3_conv1_128_32", 1, 64, 64, 128,  32, 3),
    ("d3_conv2_32_32",  1, 64, 64,  32,  32, 3),
    ("d4_conv1_32_16",  1, 64, 128, 32,  16, 3),
    ("d4_conv2_16_16",  1, 64, 128, 16,  16, 3),
    ("r50_1x1_256_64",  1, 32, 32, 256,  64, 1),
    ("r50_1x1_64_256",  1, 32, 32,  64, 256, 1),
    ("ragged_24_96_48", 3, 24, 40,  96,  48, 3),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[c[0] for c in DGRAD_CASES])
def test_native_dgrad_matches_fp32_conv_backward(case, cuda_device):
    """_ConvFn.backward's data gradient (forward kernel on gy with ``dgrad_weights``) against fp32 autograd of the same
    conv on the same bf16-rounded operands; tolerance of the per-operator tests (one bf16 rounding of the fp32 result)."""
    from unet_watermark_b200.training import _ConvFn, native_dgrad_applies
    name, n, h, w, cin, cout, k = case
    g = torch.Generator().manual_seed(len(name) * 131 + cin)
    x = torch.randn(n, cin, h, w, generator=g).to(cuda_device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cout * k * k) ** 0.5).to(cuda_device).to(torch.bfloat16).float()
    gy = torch.randn(n, cout, h, w, generator=g).to(cuda_device).to(torch.bfloat16)
    assert native_dgrad_applies(wt.shape, 1, k // 2)
    x.requires_grad_(True)
    wp = wt.clone().requires_grad_(True)
    before = _lib.load().uwm_kernel_launch_count()
    y = _ConvFn.apply(x, wp, 1, k // 2)
    y.backward(gy)
    assert _lib.load().uwm_kernel_launch_count() - before == 3           # filter pack + forward conv + data-gradient conv
    x32 = x.detach().float().requires_grad_(True)
    w32 = wt.clone().requires_grad_(True)
    torch.nn.functional.conv2d(x32, w32, padding=k // 2).backward(gy.float())
    err = (x.grad.float() - x32.grad).abs()
    assert bool((err <= 1e-2 * x32.grad.abs().clamp_min(1.0)).all()), f"dgrad max err {err.max().item()}"
    werr = (wp.grad - w32.grad).abs()
    assert bool((werr <= 2e-2 * w32.grad.abs().clamp_min(1.0)).all()), f"wgrad max err {werr.max().item()}"


def test_native_dgrad_switch_and_strided_convs_stay_on_cudnn(cuda_device, monkeypatch):
    from unet_watermark_b200.training import _ConvFn, native_dgrad_applies
    assert not native_dgrad_applies((128, 64, 3, 3), 2, 1) and not native_dgrad_applies((128, 64, 1, 1), 2, 0)
    x = torch.randn(2, 64, 16, 16, device=cuda_device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 64, 3, 3, device=cuda_device) / 24
    grads = []
    for flag in ("1", "0"):
        monkeypatch.setenv("UWM_NATIVE_DGRAD", flag)
        xi = x.clone().requires_grad_(True)
        before = _lib.load().uwm_kernel_launch_count()
        _ConvFn.apply(xi, w.clone().requires_grad_(True), 1, 1).sum().backward()
        assert _lib.load().uwm_kernel_launch_count() - before == (3 if flag == "1" else 2)   # filter pack, conv(s)
        grads.append(xi.grad.float())
    assert (grads[0] - grads[1]).abs().max() <= 2e-2 * grads[1].abs().max()


# train-mode BatchNorm (+ residual)(+ ReLU) on csrc/uwm_train.cu: (name, n, h, w, c, relu, residual)
BN_CASES = [
    ("c16_fullres",   2, 64, 128,  16, True,  False),
    ("c32",           2, 64,  64,  32, True,  False),
    ("c64_res",       2, 32,  32,  64, True,  True),
    ("c128_res",      2, 16,  16, 128, True,  True),
    ("c256_norelu",   2,  8,   8, 256, False, False),     # downsample branch: BatchNorm only
    ("c512_res",      4,  4,   4, 512, True,  True),
    ("c48_ragged",    3, 24,  40,  48, True,  False),     # 6 channel groups: 252-thread blocks
    ("c2048_res",     1,  4,   4, 2048, True, True),      # resnet50 layer4 expansion
    ("one_row_each",  1,  1,   3,  64, True,  True),      # fewer pixels than rows per block
    ("many_rows",     4, 128, 128, 64, True,  True),      # several grid-stride iterations per thread
]


@pytest.mark.parametrize("case", BN_CASES, ids=[c[0] for c in BN_CASES])
def test_native_batchnorm_matches_torch(case, cuda_device):
    """_BNFn (forward, running statistics, every gradient) against torch.nn.BatchNorm2d in fp32 on the same bf16-rounded
    operands.  Outputs are bf16: |err| <= 1e-2 * max(1, |ref|); per-channel sums are fp32-exact to ~1e-5 relative."""
    from unet_watermark_b200.training import _BNFn
    name, n, h, w, c, relu, with_res = case
    g = torch.Generator().manual_seed(c * 7 + h)
    dev = cuda_device
    x = (torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = torch.randn(n, c, h, w, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if with_res else None
    gy = torch.randn(n, c, h, w, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    ref = torch.nn.BatchNorm2d(c).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5); bn.bias.copy_(torch.randn(c, generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(c, generator=g)); bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    ref.load_state_dict(bn.state_dict())
    xi = x.clone().requires_grad_(True)
    ri = res.clone().requires_grad_(True) if with_res else None
    v0 = bn.running_mean._version
    before = _lib.load().uwm_kernel_launch_count()
    y = _BNFn.apply(xi, bn.weight, bn.bias, ri, bn, relu)
    y.backward(gy)
    assert _lib.load().uwm_kernel_launch_count() - before == 6
    assert bn.running_mean._version > v0 and int(bn.num_batches_tracked) == 1
    x32 = x.float().requires_grad_(True)
    r32 = res.float().requires_grad_(True) if with_res else None
    y32 = ref(x32)
    if with_res:
        y32 = y32 + r32
    if relu:
        y32 = y32.relu()
    # the reference masks the gradient where ITS output is positive; compare away from the bf16-rounding boundary
    y32.backward(gy.float())
    tol = lambda r: 1e-2 * r.abs().clamp_min(1.0)      # noqa: E731
    assert y.dtype == torch.bfloat16 and y.shape == x.shape
    assert bool(((y.float() - y32).abs() <= tol(y32)).all()), (y.float() - y32).abs().max().item()
    assert torch.allclose(bn.running_mean, ref.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var, ref.running_var, rtol=1e-4, atol=1e-5)
    same_mask = (y.float() > 0) == (y32 > 0) if relu else torch.ones_like(y32, dtype=torch.bool)
    assert same_mask.float().mean() > 0.999
    sums_tol = 2e-2 * (n * h * w) ** 0.5 + 1e-3                                 # bf16 masks differ on a few boundary pixels
    assert bool(((bn.bias.grad - ref.bias.grad).abs() <= 1e-3 * ref.bias.grad.abs() + sums_tol).all())
    assert bool(((bn.weight.grad - ref.weight.grad).abs() <= 1e-3 * ref.weight.grad.abs() + 3 * sums_tol).all())
    dxe = (xi.grad.float() - x32.grad).abs()
    assert bool((dxe[same_mask] <= 2e-2 * x32.grad.abs().clamp_min(1.0)[same_mask]).all()), dxe[same_mask].max().item()
    if with_res:
        dre = (ri.grad.float() - r32.grad).abs()
        assert bool((dre[same_mask] <= tol(r32.grad)[same_mask]).all())


def test_native_upcat_matches_torch(cuda_device):
    from unet_watermark_b200.training import _UpCatFn
    g = torch.Generator().manual_seed(5)
    for n, h, w, cx, cs in ((2, 8, 12, 64, 32), (1, 16, 16, 32, 0), (3, 5, 7, 16, 48)):
        x = torch.randn(n, cx, h, w, generator=g).to(cuda_device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        s = torch.randn(n, cs, 2 * h, 2 * w, generator=g).to(cuda_device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if cs else None
        gy = torch.randn(n, cx + cs, 2 * h, 2 * w, generator=g).to(cuda_device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        xi = x.clone().requires_grad_(True)
        si = s.clone().requires_grad_(True) if cs else None
        y = _UpCatFn.apply(xi, si)
        y.backward(gy)
        x32 = x.float().requires_grad_(True)
        s32 = s.float().requires_grad_(True) if cs else None
        r = torch.nn.functional.interpolate(x32, scale_factor=2, mode="nearest")
        if cs:
            r = torch.cat([r, s32], 1)
        r.backward(gy.float())
        assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y.float(), r)
        assert bool(((xi.grad.float() - x32.grad).abs() <= 1e-2 * x32.grad.abs().clamp_min(1.0)).all())
        if cs:
            assert torch.equal(si.grad.float(), s32.grad)


def test_device_prefetcher_uploads_in_order_and_reuses_its_staging(cuda_device):
    """Host batches arrive on the device intact and in order while later uploads are already in flight; the staging
    slots survive a second iteration, a shape change and a consumer that stops early."""
    from unet_watermark_b200.training import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(4, 3, 64, 64, generator=g).pin_memory(),
                (torch.rand(4, 64, 64, generator=g) > 0.5).long().pin_memory()) for _ in range(7)]
    pf = DevicePrefetcher(iter(batches), cuda_device)
    burn = torch.randn(2048, 2048, device=cuda_device)
    n = 0
    for (hx, ht), (x, t) in zip(batches, pf):
        for _ in range(4):
            burn = burn @ burn * 1e-3                              # keeps the step stream busy while the next upload runs
        assert x.device.type == "cuda" and torch.equal(x.cpu(), hx) and torch.equal(t.cpu(), ht)
        n += 1
    assert n == 7
    pf.batches = iter(batches[:3])
    for i, (x, t) in enumerate(pf):                                # stop early: the slot is released all the same
        if i == 1:
            break
    small = [(torch.randn(2, 3, 32, 32, generator=g).pin_memory(), torch.zeros(2, 32, 32, dtype=torch.long).pin_memory())
             for _ in range(3)]
    pf.batches = iter(batches[:2] + small)
    got = [(x.cpu(), t.cpu()) for x, t in pf]
    for (hx, ht), (x, t) in zip(batches[:2] + small, got):
        assert torch.equal(x, hx) and torch.equal(t, ht)
    assert list(DevicePrefetcher(iter(()), cuda_device)) == []


def test_train_step_from_prefetched_batches_equals_device_resident_batches(cuda_device):
    """lr = 0 keeps the weights fixed, so step i's loss depends on batch i alone: the prefetched batches are the right
    ones, in order, through the eager warm-up steps, the capture and the graph replays."""
    from unet_watermark_b200.training import DevicePrefetcher
    data = [synthetic_watermark_batch(4, 64, seed=700 + i, device="cpu") for i in range(5)]
    host = [(x.pin_memory(), t[:, 0].long().pin_memory()) for x, _, t in data]
    losses = []
    for use_pf in (False, True):
        torch.manual_seed(0)
        m = Unet("resnet34", encoder_weights=None).to(cuda_device)
        ts = TrainStep(m, lr=0.0, weight_decay=0.0)
        src = DevicePrefetcher(iter(host), cuda_device) if use_pf else ((x.to(cuda_device), t.to(cuda_device)) for x, t in host)
        losses.append([float(ts.step(x, t)) for x, t in src])
    # (not bit-equal: cuDNN's weight-gradient kernels and the BatchNorm sums accumulate with atomics)
    assert torch.allclose(torch.tensor(losses[0]), torch.tensor(losses[1]), rtol=2e-3), losses
    assert len(set(round(v, 3) for v in losses[0])) >= 4, losses        # the batches really differ


@pytest.mark.parametrize("shape", [(2, 16, 16, 8), (1, 6, 10, 16), (3, 32, 64, 64), (1, 2, 2, 8), (2, 64, 48, 24)],
                         ids=lambda s: "x".join(map(str, s)))
def test_native_maxpool_backward_matches_torch_including_ties(shape, cuda_device):
    """Stem max-pool of the training step: forward bit-exact with F.max_pool2d, backward equal to torch's - the inputs are
    ReLU-ed and coarsely quantised so that most windows hold ties (zeros, repeated values), which only the first-maximum
    arg-max rule routes identically."""
    from unet_watermark_b200.training import _MaxPoolFn
    n, h, w, c = shape
    g = torch.Generator().manual_seed(n * 1000 + h)
    x = (torch.randn(n, c, h, w, generator=g).clamp_min(0) * 2).round() / 2      # values in {0, 0.5, 1, ...}
    x = x.to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(n, c, h // 2, w // 2, generator=g).to(cuda_device, torch.bfloat16)
    gy = gy.contiguous(memory_format=torch.channels_last)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    yr.backward(gy)
    xo = x.clone().requires_grad_(True)
    yo = _MaxPoolFn.apply(xo)
    yo.backward(gy)
    assert torch.equal(yo, yr)
    # overlapping windows add up to four bf16 gradients in fp32, rounded once: equal up to the order of that sum
    assert torch.allclose(xo.grad.float(), xr.grad.float(), rtol=1e-2, atol=1e-6)
    assert (xo.grad == xr.grad).float().mean() > 0.999
    assert torch.equal(xo.grad == 0, xr.grad == 0)                              # same arg-max positions


@pytest.mark.parametrize("shape", [(64, 64, 3, 3), (48, 96, 3, 3), (256, 64, 1, 1), (16, 16, 3, 3), (40, 24, 3, 3),
                                   (512, 512, 3, 3), (64, 3, 7, 7), (8, 8, 2, 2)], ids=lambda s: "x".join(map(str, s)))
def test_pack_train_weights_equals_the_torch_restatement(shape, cuda_device):
    """One launch per conv and step: bf16 forward operand [Cout][taps][Cin] and data-gradient operand (training.dgrad_weights)."""
    from unet_watermark_b200 import ops
    from unet_watermark_b200.training import dgrad_weights
    w = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).to(cuda_device)
    w16 = w.to(torch.bfloat16)
    fwd, dg = ops.pack_train_weights(w, True)
    assert torch.equal(fwd, w16.permute(0, 2, 3, 1).reshape(shape[0], -1))
    assert torch.equal(dg, dgrad_weights(w16))
    fwd2, none = ops.pack_train_weights(w, False)
    assert none is None and torch.equal(fwd2, fwd)


def test_conv_fn_native_pack_switch_gives_the_same_gradients(cuda_device, monkeypatch):
    from unet_watermark_b200.training import _ConvFn
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 24, 24, generator=g).to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(48, 32, 3, 3, generator=g).to(cuda_device) * 0.05
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("UWM_NATIVE_PACK", flag)
        xi, wi = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = _ConvFn.apply(xi, wi, 1, 1)
        y.float().square().sum().backward()
        outs.append((y.detach(), xi.grad, wi.grad))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-2, atol=1e-3)      # cuDNN's weight gradient may split K with atomics


def test_copy_channels_between_slices_of_wider_buffers(cuda_device):
    from unet_watermark_b200 import ops
    g = torch.Generator().manual_seed(9)
    wide = torch.randn(2, 9, 7, 96, generator=g).to(cuda_device, torch.bfloat16)
    dense = ops.copy_channels(wide[..., 32:80])                                  # slice -> new dense tensor
    assert dense.is_contiguous() and torch.equal(dense, wide[..., 32:80])
    out = torch.zeros(2, 9, 7, 128, dtype=torch.bfloat16, device=cuda_device)
    ops.copy_channels(dense, out[..., 16:64])                                    # dense -> slice
    ops.copy_channels(wide[..., 0:16], out[..., 96:112])                         # slice -> slice
    want = torch.zeros_like(out)
    want[..., 16:64] = wide[..., 32:80]
    want[..., 96:112] = wide[..., 0:16]
    assert torch.equal(out, want)
    with pytest.raises(RuntimeError, match="copy_channels"):
        ops.copy_channels(wide[..., 4:20])                                       # base address not 16-byte aligned
