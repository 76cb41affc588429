"""Fixture (iii) of SURVEY.md App. D: a briefly trained synthetic-watermark checkpoint.

Random-init nets put a lot of probability mass at the decision threshold, so their mask agreement under
bf16 is fixture noise.  A few hundred optimisation steps on synthetic watermarks give
bimodal logits, which is what the north star's ">= 99.9 % of pixels" criterion is about.  Training uses the
ORACLE module with stock torch autograd on whatever device is given (test infrastructure, not product).
"""
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O


def synthetic_watermark_batch(batch, size, seed, device="cpu"):
    """Image-like backgrounds with 1-3 alpha-blended bright rectangles/ellipses; returns (x normalised
    fp32 [B,3,S,S], u8 NHWC, target float [B,1,S,S])."""
    g = torch.Generator().manual_seed(seed)
    base = O.image_like_input(batch, size, seed=seed, normalise=False) * 0.7      # [B,3,S,S] in [0,0.7]
    target = torch.zeros(batch, 1, size, size)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    for b in range(batch):
        for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
            cy, cx = (torch.rand(2, generator=g) * size).tolist()
            ry, rx = ((0.06 + 0.16 * torch.rand(2, generator=g)) * size).tolist()
            if torch.rand(1, generator=g).item() < 0.5:
                m = ((yy - cy).abs() < ry) & ((xx - cx).abs() < rx)
            else:
                m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
            target[b, 0][m] = 1.0
    alpha = 0.55
    stripes = (((xx + yy) // 4) % 2).float().view(1, 1, size, size) * 0.15 + 0.85  # watermark texture
    img = base * (1 - alpha * target) + alpha * target * stripes
    img = img.clamp(0, 1)
    u8 = (img * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    mean = torch.tensor(O.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD).view(1, 3, 1, 1)
    x = (u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    return x.to(device), u8.to(device), target.to(device)


def train_fixture(device, steps=200, size=128, batch=16, seed=0, encoder="resnet34"):
    """Dice+BCE (config-5 composition), Adam lr 1e-3; returns the trained oracle on CPU in eval mode."""
    torch.manual_seed(seed)
    model = O.Unet(encoder).to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    for it in range(steps):
        x, _, t = synthetic_watermark_batch(batch, size, seed=1000 + it, device=device)
        loss = O.dice_bce_loss(model(x), t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    return model.cpu().eval(), float(loss.detach())
