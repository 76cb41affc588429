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


def pad_bias(b: torch.Tensor, cout_pad: int) -> torch.Tensor:
    out = torch.zeros(cout_pad, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.detach().float()
    return out.contiguous()
