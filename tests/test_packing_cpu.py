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


def _emulate_upcat_subpixel(x, skip, wp, c_x, cout):
    """What the SPX kernel computes from UWM_PACK_UPCAT_SUBPIXEL weights, restated with torch slicing: x chunks meet
    all 9 taps of x's zero-padded halo, parity plane (ph,pw) of skip meets taps {1-ph,2-ph} x {1-pw,2-pw} of ITS halo,
    and GEMM column (qh*2+qw)*cout + co of source pixel (i,j) lands on output pixel (2i+qh, 2j+qw)."""
    n, _, h, w = x.shape
    c_s = skip.shape[1]
    acc = torch.zeros(n, 4 * cout, h, w)
    k = 0

    def mac(a_halo, r, c):
        nonlocal k
        sl = wp[:, k * 64:(k + 1) * 64]
        k += 1
        # the kernel only loads the rows tap (r,c) can reach: all (r == 1), the qh half (c == 1) or one parity group
        qn = 4 if r == 1 else (2 if c == 1 else 1)
        qoff = 0 if r == 1 else 2 * (r == 2) + (0 if c == 1 else (c == 2))
        rows = slice(qoff * cout, (qoff + qn) * cout)
        rest = torch.ones(4 * cout, dtype=torch.bool)
        rest[rows] = False
        assert not sl[rest].any(), "rows outside the tap's reach must be structurally zero"
        acc[:, rows].add_(torch.einsum("ok,nkhw->nohw", sl[rows], a_halo[:, :, r:r + h, c:c + w]))

    for ch in range(c_x // 64):
        halo = F.pad(x[:, ch * 64:(ch + 1) * 64], (1, 1, 1, 1))
        for tap in packing.SPX_X_TAP_ORDER:
            mac(halo, tap // 3, tap % 3)
    for ph in range(2):
        for pw in range(2):
            for cc in range(c_s // 64):
                halo = F.pad(skip[:, cc * 64:(cc + 1) * 64, ph::2, pw::2], (1, 1, 1, 1))
                for r in (1 - ph, 2 - ph):
                    for c in (1 - pw, 2 - pw):
                        mac(halo, r, c)
    assert k * 64 == wp.shape[1]
    y = acc.reshape(n, 2, 2, cout, h, w)                                   # [n, qh, qw, co, i, j]
    return y.permute(0, 3, 4, 1, 5, 2).reshape(n, cout, 2 * h, 2 * w)


def test_upcat_subpixel_packing_equals_conv_on_upsample_concat():
    """conv3x3(cat(nearest_up2x(x), skip)) == the sub-pixel GEMM over x's grid with parity planes of skip."""
    g = torch.Generator().manual_seed(3)
    for cout, c_x, c_s, h, w in ((16, 64, 64, 5, 6), (32, 128, 64, 3, 4)):
        wt = torch.randn(cout, c_x + c_s, 3, 3, generator=g) / ((c_x + c_s) * 9) ** 0.5
        wt = wt.to(torch.bfloat16).float()
        x = torch.randn(2, c_x, h, w, generator=g)
        skip = torch.randn(2, c_s, 2 * h, 2 * w, generator=g)
        ref = F.conv2d(torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), skip], 1), wt, padding=1)
        wp = packing.pack_upcat_subpixel(wt, c_x)
        assert wp.dtype == torch.bfloat16 and wp.shape == (4 * cout, 64 * (9 * c_x // 64 + 16 * c_s // 64))
        got = _emulate_upcat_subpixel(x, skip, wp.float(), c_x, cout)
        # only the pre-summed x taps are re-rounded to bf16 (skip weights are copied bit-exactly)
        assert torch.allclose(got, ref, atol=2e-2, rtol=0), (got - ref).abs().max()
        # skip part alone is exact up to fp32 summation order
        x0 = torch.zeros_like(x)
        ref0 = F.conv2d(torch.cat([F.interpolate(x0, scale_factor=2, mode="nearest"), skip], 1), wt, padding=1)
        assert torch.allclose(_emulate_upcat_subpixel(x0, skip, wp.float(), c_x, cout), ref0, atol=1e-4, rtol=1e-4)


def _s2d(x):
    """[N,C,2h,2w] -> [N,4C,h,w], channel (ph*2+pw)*C + c (the layout the sub-pixel conv's GEMM writes)."""
    n, c, hh, ww = x.shape
    return x.reshape(n, c, hh // 2, 2, ww // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, hh // 2, ww // 2)


def _d2s(y, c):
    n, _, h, w = y.shape
    return y.reshape(n, 2, 2, c, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c, 2 * h, 2 * w)


def test_s2d_packing_equals_conv_at_full_resolution():
    """conv3x3(x) == depth_to_space(conv3x3_over_blocks(space_to_depth(x), pack_s2d_conv3x3(w))), and the kernel only
    issues the (block tap, parity plane) pairs that can meet: every other K slice of the packed matrix must be zero."""
    g = torch.Generator().manual_seed(4)
    for cout, rows in ((16, 0), (1, 16)):
        cin = 16
        w = (torch.randn(cout, cin, 3, 3, generator=g) / 12).to(torch.bfloat16).float()
        x = torch.randn(2, cin, 10, 12, generator=g)
        ref = F.conv2d(x, w, padding=1)
        wp = packing.pack_s2d_conv3x3(w, rows)
        assert wp.shape == (max(rows, 4 * cout), 9 * 4 * cin) and wp.dtype == torch.bfloat16
        wk = wp.float().reshape(-1, 3, 3, 4 * cin).permute(0, 3, 1, 2)              # [(q,co), (p,ci), r, c]
        y = F.conv2d(_s2d(x), wk, padding=1)[:, :4 * cout]
        assert torch.allclose(_d2s(y, cout), ref, atol=1e-5, rtol=1e-5)
        # weights are copied, never summed: the set of non-zero values is exactly w's
        assert set(wp.float().unique().tolist()) <= set(w.unique().tolist()) | {0.0}
        wv = wp.float().reshape(-1, 3, 3, 2, 2, cin)                                  # [row, r, c, ph, pw, ci]
        for r in range(3):
            for c in range(3):
                for ph in range(2):
                    for pw in range(2):
                        meets = r in (1 - ph, 2 - ph) and c in (1 - pw, 2 - pw)
                        if not meets:
                            assert not wv[:, r, c, ph, pw].any(), (r, c, ph, pw)


def test_subpixel_then_s2d_chain_matches_reference_block():
    """Decoder block 4 as the plan runs it: conv1 = sub-pixel conv WITHOUT the pixel shuffle (its GEMM output is the
    space-to-depth layout), conv2 and the head consume that layout directly."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 32, 6, 8, generator=g)
    w1 = torch.randn(16, 32, 3, 3, generator=g) / 17
    w2 = torch.randn(16, 16, 3, 3, generator=g) / 12
    wh = torch.randn(1, 16, 3, 3, generator=g) / 12
    t = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w1, padding=1).relu()
    t = F.conv2d(t, w2, padding=1).relu()
    ref = F.conv2d(t, wh, padding=1)
    k1 = packing.pack_up2x_shuffle_f32(w1).reshape(64, 3, 3, 32).permute(0, 3, 1, 2)
    k2 = packing.pack_s2d_conv3x3(w2.to(torch.bfloat16).float()).float().reshape(64, 3, 3, 64).permute(0, 3, 1, 2)
    kh = packing.pack_s2d_conv3x3(wh.to(torch.bfloat16).float(), 16).float().reshape(16, 3, 3, 64).permute(0, 3, 1, 2)
    s = F.conv2d(x, k1, padding=1).relu()                      # [1, 4x16, 6, 8] == s2d of conv1's output
    assert torch.allclose(_d2s(s, 16), F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w1, padding=1).relu(),
                          atol=1e-4, rtol=1e-4)
    s = F.conv2d(s, k2, padding=1).relu()
    z = F.conv2d(s, kh, padding=1)[:, :4]
    assert torch.allclose(_d2s(z, 1), ref, atol=2e-2, rtol=0)   # w2 / wh rounded to bf16 on this side only


def test_s2_planes_packing_equals_stride2_conv():
    """conv3x3(stride 2, pad 1) == sum over parity planes of stride-1 taps on a 2x2 block halo at origin -1, with the
    slices in the order pack_s2_planes emits them (what the SPX==2 kernel issues)."""
    g = torch.Generator().manual_seed(6)
    cout, cin, h, w = 64, 128, 10, 12
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / 30).to(torch.bfloat16).float()
    x = torch.randn(2, cin, h, w, generator=g)
    ref = F.conv2d(x, wt, stride=2, padding=1)
    wp = packing.pack_s2_planes(wt).float()
    assert wp.shape == (cout, 9 * cin)
    ho, wo = h // 2, w // 2
    acc = torch.zeros(2, cout, ho, wo)
    k = 0
    for ph in range(2):
        for pw in range(2):
            for cc in range(cin // 64):
                halo = F.pad(x[:, cc * 64:(cc + 1) * 64, ph::2, pw::2], (1, 0, 1, 0))    # block -1 in front
                for r in ((1,) if ph == 0 else (0, 1)):
                    for c in ((1,) if pw == 0 else (0, 1)):
                        sl = wp[:, k * 64:(k + 1) * 64]
                        k += 1
                        acc += torch.einsum("ok,nkhw->nohw", sl, halo[:, :, r:r + ho, c:c + wo])
    assert k * 64 == wp.shape[1]
    assert torch.allclose(acc, ref, atol=1e-4, rtol=1e-4)
    # a permutation of pack_taps' columns: same multiset of values per output channel
    assert torch.equal(wp.sort(dim=1).values, packing.pack_taps(wt).float().sort(dim=1).values)


def test_two_launch_upcat_split_equals_the_single_conv():
    """Decoder blocks 0/1: conv3x3(cat(up2x(x), skip)) == conv3x3(skip; last c_skip channels, bias) + sub-pixel conv of
    the first c_x channels (no bias) - the two launches the plan issues (UWM_PACK_TAPS_SKIP_PART /
    UWM_PACK_UP2X_SHUFFLE_X_PART), each N tile of the second using only the 2x2 taps of its parity."""
    g = torch.Generator().manual_seed(8)
    cout, c_x, c_s, h, w = 8, 6, 5, 4, 5
    wt = torch.randn(cout, c_x + c_s, 3, 3, generator=g)
    b = torch.randn(cout, generator=g)
    x = torch.randn(2, c_x, h, w, generator=g)
    skip = torch.randn(2, c_s, 2 * h, 2 * w, generator=g)
    ref = F.conv2d(torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), skip], 1), wt, b, padding=1)
    part = F.conv2d(skip, wt[:, c_x:], b, padding=1)                                   # launch 1 (bias here)
    wq = packing.pack_up2x_shuffle_f32(wt[:, :c_x]).reshape(2, 2, cout, 3, 3, c_x)     # [qh, qw, co, a, b, ci]
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for qh in range(2):
        for qw in range(2):
            acc = torch.zeros(2, cout, h, w)
            for a in (qh, qh + 1):                      # the N tile of parity (qh,qw) issues taps {qh,qh+1} x {qw,qw+1}
                for bb in (qw, qw + 1):
                    acc += torch.einsum("ok,nkhw->nohw", wq[qh, qw, :, a, bb, :], xp[:, :, a:a + h, bb:bb + w])
            # ... and every other tap of that parity's weights is structurally zero
            rest = wq[qh, qw].clone()
            rest[:, qh:qh + 2, qw:qw + 2, :] = 0
            assert not rest.any()
            out[:, :, qh::2, qw::2] = acc
    assert torch.allclose(out + part, ref, atol=1e-4, rtol=1e-4)


def test_normalize_as_one_fma_equals_both_reference_operation_orders_in_bf16():
    """prep_s2d_kernel normalises a byte with ONE fused multiply-add, fma(k, 1/(255 std), -mean/std) (csrc/glue.cuh).
    For the ImageNet constants of get_val_transform (reference src/utils/dataset.py:389-395) that rounds to the same
    bf16 as albumentations' fp32 order (k - 255 mean) * (1 / (255 std)) and as (k/255 - mean)/std, for every possible
    input (256 values x 3 channels): checked here exhaustively, with the constants the kernel hard-codes."""
    import numpy as np
    na = np.array([0.017124755, 0.017507004, 0.017429195], dtype=np.float32)      # csrc/glue.cuh
    nb = np.array([-2.117904, -2.0357141, -1.8044444], dtype=np.float32)
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32)
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32)
    k = np.arange(256, dtype=np.float32)

    def bf16(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).view(torch.int16).numpy()
    for c in range(3):
        assert na[c] == np.float32(np.float32(1.0) / np.float32(255.0) / std[c]) and nb[c] == np.float32(-mean[c] / std[c])
        fma = (k.astype(np.float64) * np.float64(na[c]) + np.float64(nb[c])).astype(np.float32)     # exact product and sum in fp64 = fma
        alb = ((k - np.float32(mean[c] * np.float32(255))) * np.reciprocal(np.float32(std[c] * np.float32(255)))).astype(np.float32)
        div = (((k * (np.float32(1) / np.float32(255))).astype(np.float32) - mean[c]).astype(np.float32) / std[c]).astype(np.float32)
        assert np.array_equal(bf16(fma), bf16(alb)) and np.array_equal(bf16(fma), bf16(div))
