"""Host-side weight transformations (CPU, fp32): BN folding, tap packing, stem space-to-depth packing and the
sub-pixel packing of conv3x3(nearest_up2x(x)) are exact re-arrangements — checked against plain torch ops."""
import torch
import torch.nn.functional as F

from unet_watermark_b200 import packing


def test_fold_bn_equals_conv_then_bn():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 4, 3, 3, generator=g)
    gamma, beta = torch.rand(8, generator=g) + 0.5, torch.randn(8, generator=g)
    mean, var = torch.randn(8, generator=g), torch.rand(8, generator=g) + 0.5
    x = torch.randn(2, 4, 9, 9, generator=g)
    wf, bf = packing.fold_bn(w, gamma, beta, mean, var)
    ref = F.batch_norm(F.conv2d(x, w, padding=1), mean, var, gamma, beta, training=False, eps=1e-5)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), ref, atol=1e-5, rtol=1e-5)


def test_pack_taps_layout():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = packing.pack_taps(w, 16).float()
    assert p.shape == (16, 27) and bool((p[2:] == 0).all())
    assert p[1, (1 * 3 + 2) * 3 + 1] == w[1, 1, 1, 2]          # K index = (i*kw + j)*Cin + c


def test_stem_s2d_packing_equals_7x7_stride2_conv():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(64, 3, 7, 7, generator=g)
    x = torch.randn(1, 3, 16, 20, generator=g)
    ref = F.conv2d(x, w, stride=2, padding=3)
    # 2x2 space-to-depth of x, channel (ph*2+pw)*3 + c, padded to 16 channels
    xs = torch.zeros(1, 16, 8, 10)
    for ph in range(2):
        for pw in range(2):
            xs[:, (ph * 2 + pw) * 3:(ph * 2 + pw) * 3 + 3] = x[:, :, ph::2, pw::2]
    wp = packing.pack_stem_s2d(w).float().reshape(64, 4, 4, 16).permute(0, 3, 1, 2)   # [co, ch, r, s]
    got = F.conv2d(F.pad(xs, (2, 1, 2, 1)), wp)                                        # taps -2..1
    # the packed weights are bf16-rounded: compare against the reference on rounded weights
    ref_r = F.conv2d(x, w.to(torch.bfloat16).float(), stride=2, padding=3)
    assert torch.allclose(got, ref_r, atol=1e-4, rtol=1e-4) and ref.shape == got.shape


def test_subpixel_packing_equals_conv_on_upsampled_input():
    """conv3x3(nearest_up2x(x)) == pixel_shuffle(conv3x3(x, pre-summed weights)), exactly in fp32."""
    g = torch.Generator().manual_seed(2)
    cout, cin = 16, 32
    w = torch.randn(cout, cin, 3, 3, generator=g)
    x = torch.randn(2, cin, 6, 7, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    # fp32 version of packing.pack_up2x_shuffle (same index arithmetic, no bf16 rounding)
    out = torch.zeros(2, 2, cout, 3, 3, cin)
    for ph in range(2):
        for pw in range(2):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    out[ph, pw, :, (ph + dy) // 2 + 1, (pw + dx) // 2 + 1, :] += w[:, :, dy + 1, dx + 1]
    wc = out.reshape(4 * cout, 3, 3, cin).permute(0, 3, 1, 2)            # [(ph,pw,co), ci, a, b]
    y = F.conv2d(x, wc, padding=1).reshape(2, 2, 2, cout, 6, 7)          # [n, ph, pw, co, i, j]
    got = y.permute(0, 3, 4, 1, 5, 2).reshape(2, cout, 12, 14)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)
    # and the shipped packing is that matrix rounded once to bf16, K index (a*3+b)*cin + c
    wp = packing.pack_up2x_shuffle(w).float()
    assert wp.shape == (4 * cout, 9 * cin)
    assert torch.equal(wp, out.reshape(4 * cout, 9 * cin).to(torch.bfloat16).float())
