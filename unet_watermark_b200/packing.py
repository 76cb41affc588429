"""BatchNorm folding and weight packing for the tcgen05 conv kernel.

Eval-mode BatchNorm2d ``y = (x-mu)/sqrt(var+eps)*gamma+beta`` after a bias-free conv is folded
into the conv in fp32 (``w' = w*gamma/sqrt(var+eps)``, ``b' = beta-mu*gamma/sqrt(var+eps)``) and
the weights are rounded ONCE to bf16 (torchvision BN eps 1e-5).
"""
from __future__ import annotations

import torch


def fold_bn(w: torch.Tensor, gamma, beta, mean, var, eps: float = 1e-5):
    """[Cout,Cin,kh,kw] fp32 conv weight + BN statistics -> (folded weight, bias), fp32."""
    w = w.detach().float()
    scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
    return w * scale.view(-1, 1, 1, 1), beta.detach().float() - mean.detach().float() * scale


def pack_taps(w: torch.Tensor, cout_pad: int | None = None) -> torch.Tensor:
    """[Cout,Cin,kh,kw] -> bf16 [cout_pad, kh*kw*Cin] with K index = (i*kw+j)*Cin + c."""
    cout, cin, kh, kw = w.shape
    cout_pad = cout_pad or (cout + 15) // 16 * 16
    out = torch.zeros(cout_pad, kh * kw * cin, dtype=torch.float32, device=w.device)
    out[:cout] = w.detach().float().permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
    return out.to(torch.bfloat16).contiguous()


def pack_stem_s2d(w: torch.Tensor, cout_pad: int | None = None) -> torch.Tensor:
    """7x7/s2/p3 stem weight [Cout,3,7,7] -> bf16 [cout_pad, 4*4*16] for the 4x4/s1 conv over the
    2x2 space-to-depth input: tap (r,s), channel (ph*2+pw)*3+c holds w[:, c, 2r+ph-1, 2s+pw-1]."""
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (3, 7, 7), "stem packing expects a 3-channel 7x7 kernel"
    cout_pad = cout_pad or (cout + 15) // 16 * 16
    out = torch.zeros(cout_pad, 4, 4, 16, dtype=torch.float32, device=w.device)
    wf = w.detach().float()
    for r in range(4):
        for ph in range(2):
            i = 2 * r + ph - 1
            if not 0 <= i < 7:
                continue
            for s in range(4):
                for pw in range(2):
                    j = 2 * s + pw - 1
                    if not 0 <= j < 7:
                        continue
                    base = (ph * 2 + pw) * 3
                    out[:cout, r, s, base:base + 3] = wf[:, :, i, j]
    return out.reshape(cout_pad, 256).to(torch.bfloat16).contiguous()


def pack_up2x_shuffle_f32(w: torch.Tensor) -> torch.Tensor:
    """3x3 conv applied to a nearest-2x upsampled input, as a 3x3 conv on the source grid with 4*Cout outputs.

    Output pixel (2i+ph, 2j+pw) reads upsampled pixels (2i+ph+dy, 2j+pw+dx), dy,dx in {-1,0,1}, i.e. source pixels
    (i + floor((ph+dy)/2), j + floor((pw+dx)/2)): taps that land on the same source pixel are summed (fp32) and the
    sum is rounded once to bf16 by the caller.  Returns fp32 [4*Cout, 9*Cin]; row (ph*2+pw)*Cout + co, K index (a*3+b)*Cin + c for
    source offset (a-1, b-1).  The kernel's epilogue scatters group (ph,pw) to pixel (2i+ph, 2j+pw)."""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3), "sub-pixel packing expects a 3x3 kernel"
    wf = w.detach().float()
    out = torch.zeros(2, 2, cout, 3, 3, cin, dtype=torch.float32, device=w.device)
    for ph in range(2):
        for pw in range(2):
            for dy in (-1, 0, 1):
                a = (ph + dy) // 2          # python floor division: -1 // 2 == -1
                for dx in (-1, 0, 1):
                    b = (pw + dx) // 2
                    out[ph, pw, :, a + 1, b + 1, :] += wf[:, :, dy + 1, dx + 1]
    return out.reshape(4 * cout, 9 * cin)


def pack_up2x_shuffle(w: torch.Tensor) -> torch.Tensor:
    """bf16 [4*Cout, 9*Cin] sub-pixel weights of a 3x3 conv over a nearest-2x upsampled input (see above)."""
    return pack_up2x_shuffle_f32(w).to(torch.bfloat16).contiguous()


# x taps of the sub-pixel upcat conv in issue order: the centre tap first (it reaches every output parity, so the
# tile's first MMA writes all GEMM columns), then row-major
SPX_X_TAP_ORDER = (4, 0, 1, 2, 3, 5, 6, 7, 8)


def pack_upcat_subpixel(w: torch.Tensor, c_x: int) -> torch.Tensor:
    """conv3x3 over concat(nearest-2x(x), skip) as a sub-pixel conv on x's grid (UWM_PACK_UPCAT_SUBPIXEL).

    w: [Cout, c_x + c_skip, 3, 3].  Output pixel (2i+qh, 2j+qw) is GEMM row (i,j), column (qh*2+qw)*Cout + co.
    K is a sequence of 64-channel slices in the order the kernel issues them:
      * per 64-channel chunk of x: the 9 taps of pack_up2x_shuffle (taps of the upsampled conv that land on the same
        source pixel summed in fp32, rounded once), centre tap first (SPX_X_TAP_ORDER);
      * per parity plane (ph,pw) of skip, per 64-channel chunk: taps r in {1-ph, 2-ph}, c in {1-pw, 2-pw} of the 3x3
        block neighbourhood.  Block offset r-1 of plane ph is skip row 2(i+r-1)+ph, which output row 2i+qh reads
        with kernel row 2(r-1)+ph-qh+1 - a zero slice entry when that is outside 0..2.
    Returns bf16 [4*Cout, 64 * (9*c_x/64 + 16*c_skip/64)]."""
    cout, cin, kh, kw = w.shape
    c_s = cin - c_x
    assert (kh, kw) == (3, 3) and c_x % 64 == 0 and c_s % 64 == 0 and c_x > 0 and c_s > 0
    wf = w.detach().float()
    slices = []
    wx = pack_up2x_shuffle_f32(wf[:, :c_x]).reshape(4 * cout, 9, c_x)
    for ch in range(c_x // 64):
        for tap in SPX_X_TAP_ORDER:
            slices.append(wx[:, tap, ch * 64:(ch + 1) * 64])
    for ph in range(2):
        for pw in range(2):
            for cc in range(c_s // 64):
                wsk = wf[:, c_x + cc * 64: c_x + (cc + 1) * 64]              # [cout, 64, 3, 3]
                for r in (1 - ph, 2 - ph):
                    for c in (1 - pw, 2 - pw):
                        sl = torch.zeros(2, 2, cout, 64, dtype=torch.float32, device=w.device)
                        for qh in range(2):
                            kr = 2 * (r - 1) + ph - qh + 1
                            for qw in range(2):
                                kcol = 2 * (c - 1) + pw - qw + 1
                                if 0 <= kr <= 2 and 0 <= kcol <= 2:
                                    sl[qh, qw] = wsk[:, :, kr, kcol]
                        slices.append(sl.reshape(4 * cout, 64))
    return torch.cat(slices, dim=1).to(torch.bfloat16).contiguous()


def pack_s2_planes(w: torch.Tensor) -> torch.Tensor:
    """Stride-2 3x3 conv (pad 1) as stride-1 taps on the input's four parity planes (UWM_PACK_S2_PLANES).

    Output pixel (i,j) reads input row 2i+kr-1: kr = 1 is (block i, plane 0), kr = 0 / 2 are (block i-1 / i, plane 1).
    With a 2x2 block halo at origin -1 (halo row r = 0 is block i-1, r = 1 block i) plane ph meets rows {1} (ph = 0)
    or {0, 1} (ph = 1); columns likewise.  Returns bf16 [Cout, 9*Cin]: 64-channel slices in issue order - plane
    (ph,pw) major, then 64-channel chunk, then taps (r,c) row-major."""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3) and cin % 64 == 0
    wb = w.detach().float()
    slices = []
    for ph in range(2):
        for pw in range(2):
            for cc in range(cin // 64):
                for r in ((1,) if ph == 0 else (0, 1)):
                    kr = 1 if ph == 0 else (0 if r == 0 else 2)
                    for c in ((1,) if pw == 0 else (0, 1)):
                        kcol = 1 if pw == 0 else (0 if c == 0 else 2)
                        slices.append(wb[:, cc * 64:(cc + 1) * 64, kr, kcol])
    return torch.cat(slices, dim=1).to(torch.bfloat16).contiguous()


def pack_s2d_conv3x3(w: torch.Tensor, rows: int = 0) -> torch.Tensor:
    """conv3x3 on a tensor stored space-to-depth (UWM_PACK_S2D_CONV).

    The [.,2h,2w,Cin] input is kept as [.,h,w,4*Cin] with channel (ph*2+pw)*Cin + ci, the output likewise with row
    (qh*2+qw)*Cout + co.  Output row 2i+qh reads input row 2i+qh+kr-1 = 2(i+r-1)+ph for block tap r, i.e. kernel row
    kr = 2(r-1)+ph-qh+1 (zero when outside 0..2; same for columns).  Returns bf16 [max(rows, 4*Cout), 9*4*Cin] in
    pack_taps order over blocks: K index (r*3+c)*4*Cin + (ph*2+pw)*Cin + ci.  Every weight of w appears in exactly
    the slots whose (block tap, plane) pair meets - nothing is summed, so the values are bit-identical to pack_taps."""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3)
    wf = w.detach().float()
    out = torch.zeros(2, 2, cout, 3, 3, 2, 2, cin, dtype=torch.float32, device=w.device)
    for qh in range(2):
        for qw in range(2):
            for r in range(3):
                for c in range(3):
                    for ph in range(2):
                        kr = 2 * (r - 1) + ph - qh + 1
                        if not 0 <= kr <= 2:
                            continue
                        for pw in range(2):
                            kc = 2 * (c - 1) + pw - qw + 1
                            if 0 <= kc <= 2:
                                out[qh, qw, :, r, c, ph, pw, :] = wf[:, :, kr, kc]
    out = out.reshape(4 * cout, 9 * 4 * cin)
    if rows > out.shape[0]:
        out = torch.cat([out, torch.zeros(rows - out.shape[0], out.shape[1], device=w.device)], 0)
    return out.to(torch.bfloat16).contiguous()


def pad_bias(b: torch.Tensor, cout_pad: int) -> torch.Tensor:
    out = torch.zeros(cout_pad, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.detach().float()
    return out.contiguous()
