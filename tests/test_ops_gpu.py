"""Per-operator parity on the GPU, through the C ABI, against a plain PyTorch fp32 reference of the same
op on the same bf16-rounded inputs.  Tolerance: |err| <= 1e-2 * max(1, |ref|) (one bf16 rounding of the
fp32 result, 2^-8 relative, plus fp32 summation-order noise)."""
import pytest
import torch
import torch.nn.functional as F

import zlib

from tests.op_cases import (CONV_CASES, S2D_CASES, S2P_CASES, SHUFFLE_CASES, SHUFFLE_RES_CASES, SPX_CASES,
                            UPCAT_CASES)
from unet_watermark_b200 import _lib, ops, packing


def _seed(name: str) -> int:
    """Per-case seed that is the same in every process (str hashes are salted per interpreter run)."""
    return zlib.crc32(name.encode()) % 1000

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_matches_fp32_reference(case, cuda_device):
    name, n, h, w, cin, cout, k, stride, pad, relu, use_res, in_extra, out_extra = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    xbuf = torch.randn(n, h, w, cin + in_extra, generator=g).to(dev).to(torch.bfloat16)
    x = xbuf[..., in_extra:] if in_extra else xbuf
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    res = torch.randn(n, ho, wo, cout, generator=g).to(dev).to(torch.bfloat16) if use_res else None
    wp = packing.pack_taps(wt)
    obuf = torch.full((n, ho, wo, cout + out_extra), 7.0, dtype=torch.bfloat16, device=dev)
    out = obuf[..., out_extra:] if out_extra else obuf
    before = _lib.load().uwm_kernel_launch_count()
    ops.conv2d(x, wp, bias, k, k, stride, pad, relu=relu, residual=res, out=out)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(cout, k, k, cin).permute(0, 3, 1, 2), bias,
                   stride=stride, padding=pad)
    if res is not None:
        ref = ref + res.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 * ref.abs().clamp_min(1.0)).all()), f"max err {err.max().item()}"
    if out_extra:
        assert bool((obuf[..., :out_extra] == 7.0).all()), "conv wrote outside its channel slice"


@pytest.mark.parametrize("case", UPCAT_CASES, ids=[c[0] for c in UPCAT_CASES])
def test_upcat_conv_matches_fp32_reference(case, cuda_device):
    """interpolate(nearest, x2) + cat + conv3x3 + bias + ReLU in one kernel vs the same three torch ops."""
    name, n, h, w, cx, cs, cout, up, relu, x_extra, s_extra = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    xbuf = torch.randn(n, h, w, cx + x_extra, generator=g).to(dev).to(torch.bfloat16)
    x = xbuf[..., x_extra:] if x_extra else xbuf
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    skip = None
    if cs:
        sbuf = torch.randn(n, ho, wo, cs + s_extra, generator=g).to(dev).to(torch.bfloat16)
        skip = sbuf[..., :cs] if s_extra else sbuf
    cin = cx + cs
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    wp = packing.pack_taps(wt)
    before = _lib.load().uwm_kernel_launch_count()
    out = ops.conv2d_upcat(x, skip, wp, bias, relu=relu, upsample=up)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    xi = x.float().permute(0, 3, 1, 2)
    if up:
        xi = F.interpolate(xi, scale_factor=2, mode="nearest")
    if skip is not None:
        xi = torch.cat([xi, skip.float().permute(0, 3, 1, 2)], dim=1)
    ref = F.conv2d(xi, wp.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), bias, padding=1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 * ref.abs().clamp_min(1.0)).all()), f"max err {err.max().item()}"


@pytest.mark.parametrize("case", SHUFFLE_CASES, ids=[c[0] for c in SHUFFLE_CASES])
def test_subpixel_conv_matches_fp32_reference(case, cuda_device):
    """conv3x3(interpolate(x, 2, nearest)) computed on the source grid (pre-summed taps + pixel shuffle) vs the two
    torch ops.  The pre-summed weights are rounded once to bf16, so the tolerance is the per-op one plus one extra
    weight rounding: |err| <= 1.5e-2 * max(1, |ref|) against the reference on the ORIGINAL bf16-rounded weights."""
    name, n, h, w, cin, cout, relu = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    ws = packing.pack_up2x_shuffle(wt)
    before = _lib.load().uwm_kernel_launch_count()
    out = ops.conv2d_up2x_shuffle(x, ws, bias.repeat(4).contiguous(), relu=relu)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    assert out.shape == (n, 2 * h, 2 * w, cout)
    xi = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(xi, wt.to(torch.bfloat16).float(), bias, padding=1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert bool((err <= 1.5e-2 * ref.abs().clamp_min(1.0)).all()), f"max err {err.max().item()}"


@pytest.mark.parametrize("case", SPX_CASES, ids=[c[0] for c in SPX_CASES])
def test_upcat_subpixel_conv_matches_fp32_reference(case, cuda_device):
    """conv3x3(cat(interpolate(x, 2, nearest), skip)) computed on x's grid (x taps pre-summed, skip read as four parity
    planes with 2x2 taps each, pixel-shuffle epilogue) vs the torch ops on the ORIGINAL bf16-rounded weights.
    Tolerance: the per-op one plus one extra weight rounding of the pre-summed x taps, 1.5e-2 * max(1, |ref|)."""
    name, n, h, w, cx, cs, cout, relu = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    x = torch.randn(n, h, w, cx, generator=g).to(dev).to(torch.bfloat16)
    skip = torch.randn(n, 2 * h, 2 * w, cs, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cx + cs, 3, 3, generator=g) / ((cx + cs) * 9) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    ws = packing.pack_upcat_subpixel(wt, cx)
    before = _lib.load().uwm_kernel_launch_count()
    out = ops.conv2d_upcat_subpixel(x, skip, ws, bias.repeat(4).contiguous(), relu=relu)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    assert out.shape == (n, 2 * h, 2 * w, cout)
    xi = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(torch.cat([xi, skip.float().permute(0, 3, 1, 2)], 1), wt.to(torch.bfloat16).float(), bias, padding=1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert bool((err <= 1.5e-2 * ref.abs().clamp_min(1.0)).all()), f"max err {err.max().item()}"


@pytest.mark.parametrize("case", S2P_CASES, ids=[c[0] for c in S2P_CASES])
def test_stride2_plane_conv_matches_fp32_reference(case, cuda_device):
    """Stride-2 conv3x3 over the input's parity planes vs F.conv2d(stride=2) on the same bf16 inputs / weights, and
    against the library's other stride-2 kernel (conv_tc) on the same data: |err| <= 1e-2 * max(1, |ref|)."""
    name, n, h, w, cin, cout, relu = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    before = _lib.load().uwm_kernel_launch_count()
    out = ops.conv2d_s2_planes(x, packing.pack_s2_planes(wt), bias, relu=relu)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    assert out.shape == (n, h // 2, w // 2, cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, stride=2, padding=1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 * ref.abs().clamp_min(1.0)).all()), f"max err {err.max().item()}"
    other = ops.conv2d(x, packing.pack_taps(wt), bias, 3, 3, 2, 1, relu=relu)
    assert bool(((out.float() - other.float()).abs() <= 2e-2 * ref.abs().clamp_min(1.0)).all())


def _s2d_nhwc(x):
    """[N,2h,2w,C] -> [N,h,w,4C], channel (ph*2+pw)*C + c."""
    n, hh, ww, c = x.shape
    return x.reshape(n, hh // 2, 2, ww // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, hh // 2, ww // 2, 4 * c).contiguous()


@pytest.mark.parametrize("case", S2D_CASES, ids=[c[0] for c in S2D_CASES])
def test_s2d_conv_matches_fp32_reference(case, cuda_device):
    """conv3x3 on a 16-channel tensor stored space-to-depth vs F.conv2d at full resolution on the same bf16 inputs and
    weights (the packing copies weights, so only fp32 accumulation order and the bf16 output rounding differ:
    |err| <= 1e-2 * max(1, |ref|), the per-op tolerance of the other conv tests)."""
    name, n, h, w, relu = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    x = torch.randn(n, 2 * h, 2 * w, 16, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(16, 16, 3, 3, generator=g) / 12).to(dev)
    bias = torch.randn(16, generator=g).to(dev)
    before = _lib.load().uwm_kernel_launch_count()
    out = ops.conv2d_s2d(_s2d_nhwc(x), packing.pack_s2d_conv3x3(wt), bias.repeat(4).contiguous(), relu=relu)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, padding=1)
    if relu:
        ref = ref.relu()
    ref = _s2d_nhwc(ref.permute(0, 2, 3, 1).contiguous())
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 * ref.abs().clamp_min(1.0)).all()), f"max err {err.max().item()}"


@pytest.mark.parametrize("case", S2D_CASES, ids=[c[0] for c in S2D_CASES])
def test_s2d_head_matches_fp32_reference(case, cuda_device):
    """Head on the space-to-depth tensor: fp32 logits within 2e-3 of F.conv2d on the same bf16 inputs / weights, the
    uint8 mask bit-exact wherever the reference logit is further than that from the threshold."""
    name, n, h, w, _ = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name) + 1)
    x = torch.randn(n, 2 * h, 2 * w, 16, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(1, 16, 3, 3, generator=g) / 12).to(dev)
    bias = torch.randn(1, generator=g).to(dev)
    logits, mask = ops.head_s2d(_s2d_nhwc(x), packing.pack_s2d_conv3x3(wt, 16), packing.pad_bias(bias, 16), threshold=0.5)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, padding=1)[:, 0]
    assert logits.shape == ref.shape and mask.shape == ref.shape
    assert (logits - ref).abs().max().item() <= 2e-3
    sure = ref.abs() > 2e-3
    assert torch.equal(mask[sure], ((ref > 0).to(torch.uint8) * 255)[sure])
    assert set(mask.unique().tolist()) <= {0, 255}
    # mask-only call (what predict_mask uses) gives the same bytes
    _, mask2 = ops.head_s2d(_s2d_nhwc(x), packing.pack_s2d_conv3x3(wt, 16), packing.pad_bias(bias, 16), threshold=0.5,
                            want_logits=False)
    assert torch.equal(mask, mask2)


@pytest.mark.parametrize("case", SHUFFLE_RES_CASES, ids=[c[0] for c in SHUFFLE_RES_CASES])
def test_subpixel_parity_tiles_with_residual(case, cuda_device):
    """act(conv3x3(interpolate(x, 2, nearest)) + bias + residual) with Cout 128 / 256: every output parity is its own N
    tile and issues 4 of the 9 taps.  Same tolerance as the other sub-pixel conv (pre-summed weights rounded once)."""
    name, n, h, w, cin, cout, relu, with_res = case
    dev = cuda_device
    g = torch.Generator().manual_seed(_seed(name))
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(n, 2 * h, 2 * w, cout, generator=g).to(dev).to(torch.bfloat16) if with_res else None
    before = _lib.load().uwm_kernel_launch_count()
    out = ops.conv2d_up2x_shuffle_res(x, packing.pack_up2x_shuffle(wt), bias.repeat(4).contiguous(), residual=res, relu=relu)
    assert _lib.load().uwm_kernel_launch_count() == before + 1
    xi = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    conv = F.conv2d(xi, wt.to(torch.bfloat16).float(), bias, padding=1)
    ref = conv + res.float().permute(0, 3, 1, 2) if res is not None else conv
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    # the conv term carries the extra weight rounding; where it cancels against the residual the bound must follow the
    # conv's magnitude, not the (small) sum's
    scale = torch.maximum(ref.abs(), conv.abs().permute(0, 2, 3, 1)).clamp_min(1.0)
    err = (out.float() - ref).abs()
    assert bool((err <= 1.5e-2 * scale).all()), f"max err {err.max().item()}"


def test_conv_rejects_bad_arguments(cuda_device):
    x = torch.zeros(1, 8, 8, 24, dtype=torch.bfloat16, device=cuda_device)      # cin not a multiple of 16
    w = torch.zeros(16, 9 * 24, dtype=torch.bfloat16, device=cuda_device)
    b = torch.zeros(16, device=cuda_device)
    with pytest.raises(RuntimeError, match="multiple of 16"):
        ops.conv2d(x, w, b, 3, 3, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.conv2d(x.cpu(), w.cpu(), b.cpu(), 3, 3, 1, 1)


def test_maxpool_exact(cuda_device):
    g = torch.Generator().manual_seed(1)
    # 64-channel multiples take the TMA-staged kernel (partial tiles, several chunks, several tiles per CTA), the rest
    # the register kernel; randn inputs are negative at the borders, so zero padding instead of -inf would show
    for shape in ((2, 32, 48, 64), (1, 6, 10, 8), (3, 2, 2, 128), (1, 36, 52, 64), (5, 128, 160, 64), (2, 20, 12, 192)):
        x = torch.randn(*shape, generator=g).to(cuda_device).to(torch.bfloat16)
        ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        assert torch.equal(ops.maxpool3x3s2(x).float(), ref)
    # pitched input (the stem feature lives inside a concat buffer)
    buf = torch.randn(2, 16, 16, 96, generator=g).to(cuda_device).to(torch.bfloat16)
    ref = F.max_pool2d(buf[..., 32:].float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(ops.maxpool3x3s2(buf[..., 32:]).float(), ref)


def test_upsample_into_concat_slice_exact(cuda_device):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 12, 20, 64, generator=g).to(cuda_device).to(torch.bfloat16)
    cat = torch.zeros(2, 24, 40, 96, dtype=torch.bfloat16, device=cuda_device)
    ops.upsample2x(x, out=cat[..., :64])
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(cat[..., :64].float(), ref)
    assert bool((cat[..., 64:] == 0).all())


def test_prep_input_f32_and_u8(cuda_device):
    g = torch.Generator().manual_seed(3)
    xin = torch.randn(2, 3, 64, 96, generator=g).to(cuda_device)
    xs = ops.prep_input(xin)
    ref = torch.zeros(2, 32, 48, 16, device=cuda_device)
    for ph in range(2):
        for pw in range(2):
            ref[..., (ph * 2 + pw) * 3:(ph * 2 + pw) * 3 + 3] = xin[:, :, ph::2, pw::2].permute(0, 2, 3, 1)
    assert torch.equal(xs.float(), ref.to(torch.bfloat16).float())
    u8 = torch.randint(0, 256, (2, 64, 96, 3), generator=g, dtype=torch.uint8).to(cuda_device)
    mean = torch.tensor([0.485, 0.456, 0.406], device=cuda_device)
    std = torch.tensor([0.229, 0.224, 0.225], device=cuda_device)
    xn = ((u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()
    # fused Normalize: equal to the fp32 normalisation up to one bf16 ulp (2^-8 relative)
    d = (ops.prep_input(u8).float() - ops.prep_input(xn).float()).abs()
    assert bool((d <= 2 ** -7 * xn.abs().max()).all())


def test_prep_u8_is_bit_exact_with_the_reference_normalize(cuda_device):
    """Every byte value in every channel: the fused normalisation equals bf16(albumentations' fp32 Normalize)."""
    import numpy as np
    from oracle import unet_oracle as O
    img = np.zeros((32, 32, 3), np.uint8)
    img[:16].reshape(-1, 3)[:256] = np.arange(256, dtype=np.uint8)[:, None]          # 256 values in all three channels
    img[16:] = np.random.default_rng(0).integers(0, 256, (16, 32, 3), dtype=np.uint8)
    want = O.val_transform(img, 32)                                                   # fp32 [3,32,32], no resize at 32x32
    got = ops.prep_input(torch.from_numpy(img)[None].to(cuda_device))[0].cpu()       # [16,16,16] space-to-depth
    for ph in range(2):
        for pw in range(2):
            for c in range(3):
                g = got[:, :, (ph * 2 + pw) * 3 + c]
                w = want[c, ph::2, pw::2].to(torch.bfloat16)
                assert torch.equal(g, w), (ph, pw, c)
    assert bool((got[:, :, 12:] == 0).all())


def test_stem_as_s2d_conv(cuda_device):
    g = torch.Generator().manual_seed(4)
    xin = torch.randn(2, 3, 64, 96, generator=g).to(cuda_device)
    wt = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).to(cuda_device)
    b = torch.randn(64, generator=g).to(cuda_device)
    y = ops.conv2d(ops.prep_input(xin), packing.pack_stem_s2d(wt), b, 4, 4, 1, 2, relu=True)[:, :32, :48]
    ref = F.conv2d(xin.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), b, stride=2, padding=3).relu()
    ref = ref.permute(0, 2, 3, 1)
    assert bool(((y.float() - ref).abs() <= 1e-2 * ref.abs().clamp_min(1.0)).all())


@pytest.mark.parametrize("sigmoid_out", [False, True])
def test_head_logits_mask_and_threshold_conventions(cuda_device, sigmoid_out):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 96, 16, generator=g).to(cuda_device).to(torch.bfloat16)
    wh = (torch.randn(1, 16, 3, 3, generator=g) / 12.0).to(cuda_device)
    bh = torch.tensor([0.1], device=cuda_device)
    wp = packing.pack_taps(wh)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float()[:1].view(1, 3, 3, 16).permute(0, 3, 1, 2), bh, padding=1)[:, 0]
    for thr, on_logits in ((0.5, False), (0.5, True), (0.3, False)):
        out, mask = ops.head(x, wp, packing.pad_bias(bh, 16), threshold=thr, thr_on_logits=on_logits,
                             apply_sigmoid=sigmoid_out)
        want = torch.sigmoid(ref) if sigmoid_out else ref
        assert (out - want).abs().max() < 1e-3
        cut = thr if on_logits else ops.logit(thr)
        mref = (ref > cut).to(torch.uint8) * 255
        wrong = (mask != mref) & ((ref - cut).abs() > 1e-3)       # only ties within fp32 noise may differ
        assert int(wrong.sum()) == 0
        assert set(mask.unique().tolist()) <= {0, 255}
