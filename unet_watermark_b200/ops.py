"""Single-operator wrappers over the C ABI (torch tensors in, torch tensors out).

torch is used only to own device memory and the stream; all arithmetic happens in
libuwm_b200.so.  Activations are NHWC bf16 (``[N,H,W,C]`` contiguous, or a channel slice of a
wider buffer described by its pixel pitch).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("unet_watermark_b200 kernels need CUDA tensors (no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _pitch(t: torch.Tensor) -> int:
    """pixel pitch of an NHWC view whose channel dim may be a slice of a wider buffer"""
    assert t.dim() == 4 and t.stride(3) == 1, "expected NHWC with unit channel stride"
    p = t.stride(2)
    assert t.stride(1) == p * t.shape[2] and t.stride(0) == p * t.shape[2] * t.shape[1], \
        "NHWC view must be dense in N,H,W"
    return p


def conv2d(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, kh: int, kw: int, stride: int,
           pad: int, relu: bool = False, residual: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NHWC bf16 conv (+bias)(+residual)(+ReLU).  w_packed: [Cout, kh*kw*Cin] bf16 (packing.pack_taps)."""
    _require_cuda(x, w_packed, bias, residual, out)
    lib = _lib.load()
    n, h, w, cin = x.shape
    cout = w_packed.shape[0]
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_nhwc_bf16(
        x.data_ptr(), n, h, w, cin, _pitch(x), w_packed.data_ptr(), bias.data_ptr(), cout, kh, kw, stride, pad,
        residual.data_ptr() if residual is not None else None,
        _pitch(residual) if residual is not None else 0, int(relu), out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_nhwc_bf16")
    return out


def conv2d_upcat(x: torch.Tensor, skip: Optional[torch.Tensor], w_packed: torch.Tensor, bias: torch.Tensor,
                 kh: int = 3, kw: int = 3, pad: int = 1, relu: bool = True, upsample: bool = True,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv(concat(nearest_up2x(x), skip)) (+bias)(+ReLU) in one kernel (smp DecoderBlock: interpolate+cat+conv1)."""
    _require_cuda(x, skip, w_packed, bias, out)
    lib = _lib.load()
    n, h, w, cx = x.shape
    cout = w_packed.shape[0]
    ho, wo = (2 * h, 2 * w) if upsample else (h, w)
    if skip is not None:
        assert skip.shape[:3] == (n, ho, wo), (skip.shape, (n, ho, wo))
    if out is None:
        out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_upcat_nhwc_bf16(
        x.data_ptr(), n, h, w, cx, _pitch(x), int(upsample),
        skip.data_ptr() if skip is not None else None, skip.shape[3] if skip is not None else 0,
        _pitch(skip) if skip is not None else 0, w_packed.data_ptr(), bias.data_ptr(), cout, kh, kw, pad,
        int(relu), out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_upcat_nhwc_bf16")
    return out


def conv2d_up2x_shuffle(x: torch.Tensor, w_shuffle: torch.Tensor, bias4: torch.Tensor, relu: bool = True,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv3x3(nearest_up2x(x)) (+bias)(+ReLU) as a sub-pixel conv on the source grid.
    w_shuffle: packing.pack_up2x_shuffle(w) = [4*Cout, 9*Cin] bf16; bias4: bias repeated 4x (fp32)."""
    _require_cuda(x, w_shuffle, bias4, out)
    lib = _lib.load()
    n, h, w, cin = x.shape
    cout = w_shuffle.shape[0] // 4
    if out is None:
        out = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_up2x_shuffle_nhwc_bf16(x.data_ptr(), n, h, w, cin, _pitch(x), w_shuffle.data_ptr(),
                                               bias4.data_ptr(), cout, int(relu), out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_up2x_shuffle_nhwc_bf16")
    return out


def conv2d_up2x_shuffle_res(x: torch.Tensor, w_shuffle: torch.Tensor, bias4: torch.Tensor,
                            residual: Optional[torch.Tensor] = None, relu: bool = True,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(conv3x3(nearest_up2x(x)) + bias + residual) as a sub-pixel conv; Cout up to 64, or 128 / 256 (one N tile per
    output parity).  residual: [N,2h,2w,Cout] bf16 or None."""
    _require_cuda(x, w_shuffle, bias4, residual, out)
    lib = _lib.load()
    n, h, w, cin = x.shape
    cout = w_shuffle.shape[0] // 4
    if out is None:
        out = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_up2x_shuffle_res_nhwc_bf16(x.data_ptr(), n, h, w, cin, _pitch(x), w_shuffle.data_ptr(),
                                                   bias4.data_ptr(), cout,
                                                   residual.data_ptr() if residual is not None else None,
                                                   _pitch(residual) if residual is not None else 0, int(relu),
                                                   out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_up2x_shuffle_res_nhwc_bf16")
    return out


def conv2d_upcat_subpixel(x: torch.Tensor, skip: torch.Tensor, w_spx: torch.Tensor, bias4: torch.Tensor,
                          relu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv3x3(concat(nearest_up2x(x), skip)) (+bias)(+ReLU) as a sub-pixel conv on x's grid.
    x: [N,h,w,Cx]; skip: [N,2h,2w,Cs]; w_spx: packing.pack_upcat_subpixel(w, Cx); bias4: bias repeated 4x (fp32)."""
    _require_cuda(x, skip, w_spx, bias4, out)
    lib = _lib.load()
    n, h, w, cx = x.shape
    if skip.shape[0] != n or skip.shape[1] != 2 * h or skip.shape[2] != 2 * w:
        raise ValueError(f"skip {tuple(skip.shape)} must be [N, 2h, 2w, Cs] for x {tuple(x.shape)}")
    cout = w_spx.shape[0] // 4
    if out is None:
        out = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_upcat_subpixel_nhwc_bf16(x.data_ptr(), n, h, w, cx, _pitch(x), skip.data_ptr(), skip.shape[3],
                                                 _pitch(skip), w_spx.data_ptr(), bias4.data_ptr(), cout, int(relu),
                                                 out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_upcat_subpixel_nhwc_bf16")
    return out


def head(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, threshold: Optional[float] = 0.5,
         thr_on_logits: bool = False, want_logits: bool = True, apply_sigmoid: bool = False):
    """conv3x3(Cin->1)+bias -> (fp32 logits [N,H,W] or None, uint8 mask [N,H,W] or None)."""
    _require_cuda(x, w_packed, bias)
    lib = _lib.load()
    n, h, w, cin = x.shape
    logits = torch.empty(n, h, w, dtype=torch.float32, device=x.device) if want_logits else None
    mask = torch.empty(n, h, w, dtype=torch.uint8, device=x.device) if threshold is not None else None
    thr_logit = 0.0
    if threshold is not None:
        thr_logit = float(threshold) if thr_on_logits else logit(threshold)
    rc = lib.uwm_head_nhwc_bf16(x.data_ptr(), n, h, w, cin, _pitch(x), w_packed.data_ptr(), bias.data_ptr(),
                                logits.data_ptr() if logits is not None else None, int(apply_sigmoid),
                                mask.data_ptr() if mask is not None else None, thr_logit, _stream())
    _lib.check(rc, "uwm_head_nhwc_bf16")
    return logits, mask


def conv2d_s2_planes(x: torch.Tensor, w_planes: torch.Tensor, bias: torch.Tensor, relu: bool = True,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Stride-2 conv3x3 (pad 1) on the parity-plane halo kernel.  x: [N,h,w,Cin] (h, w even, Cin % 64 == 0);
    w_planes: packing.pack_s2_planes(w) = [Cout, 9*Cin] bf16 (Cout % 64 == 0)."""
    _require_cuda(x, w_planes, bias, out)
    lib = _lib.load()
    n, h, w, cin = x.shape
    cout = w_planes.shape[0]
    if out is None:
        out = torch.empty(n, h // 2, w // 2, cout, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_s2_planes_nhwc_bf16(x.data_ptr(), n, h, w, cin, _pitch(x), w_planes.data_ptr(), bias.data_ptr(),
                                            cout, int(relu), out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_s2_planes_nhwc_bf16")
    return out


def conv2d_s2d(x: torch.Tensor, w_s2d: torch.Tensor, bias4: torch.Tensor, relu: bool = True,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv3x3 'same' on a 16-channel tensor kept space-to-depth: x, out = [N,h,w,64] standing for [N,2h,2w,16].
    w_s2d: packing.pack_s2d_conv3x3(w) = [64, 9*64] bf16; bias4: bias repeated 4x (fp32)."""
    _require_cuda(x, w_s2d, bias4, out)
    lib = _lib.load()
    n, h, w, c = x.shape
    if c != 64 or tuple(w_s2d.shape) != (64, 9 * 64):
        raise ValueError(f"conv2d_s2d needs x [N,h,w,64] and weights [64, 576], got {tuple(x.shape)}, {tuple(w_s2d.shape)}")
    if out is None:
        out = torch.empty(n, h, w, 64, dtype=torch.bfloat16, device=x.device)
    rc = lib.uwm_conv2d_s2d_nhwc_bf16(x.data_ptr(), n, h, w, _pitch(x), w_s2d.data_ptr(), bias4.data_ptr(), int(relu),
                                      out.data_ptr(), _pitch(out), _stream())
    _lib.check(rc, "uwm_conv2d_s2d_nhwc_bf16")
    return out


def head_s2d(x: torch.Tensor, w_s2d: torch.Tensor, bias: torch.Tensor, threshold: Optional[float] = 0.5,
             thr_on_logits: bool = False, want_logits: bool = True, apply_sigmoid: bool = False):
    """The head on a space-to-depth tensor x [N,h,w,64] -> (fp32 logits [N,2h,2w] or None, uint8 mask or None).
    w_s2d: packing.pack_s2d_conv3x3(w_head, 16) = [16, 9*64] bf16."""
    _require_cuda(x, w_s2d, bias)
    lib = _lib.load()
    n, h, w, c = x.shape
    if c != 64 or tuple(w_s2d.shape) != (16, 9 * 64):
        raise ValueError(f"head_s2d needs x [N,h,w,64] and weights [16, 576], got {tuple(x.shape)}, {tuple(w_s2d.shape)}")
    logits = torch.empty(n, 2 * h, 2 * w, dtype=torch.float32, device=x.device) if want_logits else None
    mask = torch.empty(n, 2 * h, 2 * w, dtype=torch.uint8, device=x.device) if threshold is not None else None
    thr_logit = 0.0
    if threshold is not None:
        thr_logit = float(threshold) if thr_on_logits else logit(threshold)
    rc = lib.uwm_head_s2d_nhwc_bf16(x.data_ptr(), n, h, w, _pitch(x), w_s2d.data_ptr(), bias.data_ptr(),
                                    logits.data_ptr() if logits is not None else None, int(apply_sigmoid),
                                    mask.data_ptr() if mask is not None else None, thr_logit, _stream())
    _lib.check(rc, "uwm_head_s2d_nhwc_bf16")
    return logits, mask


def logit(p: float) -> float:
    """threshold on sigmoid(z) > p  <=>  z > log(p/(1-p))"""
    if p <= 0.0:
        return -math.inf
    if p >= 1.0:
        return math.inf
    return math.log(p / (1.0 - p))


def maxpool3x3s2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, out)
    lib = _lib.load()
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.uwm_maxpool3x3s2_nhwc_bf16(x.data_ptr(), n, h, w, c, _pitch(x), out.data_ptr(), _pitch(out),
                                              _stream()), "uwm_maxpool3x3s2_nhwc_bf16")
    return out


def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nearest 2x; `out` may be the leading channel slice of a concat buffer"""
    _require_cuda(x, out)
    lib = _lib.load()
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.uwm_upsample2x_nhwc_bf16(x.data_ptr(), n, h, w, c, _pitch(x), out.data_ptr(), _pitch(out),
                                            _stream()), "uwm_upsample2x_nhwc_bf16")
    return out


def prep_input(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW (normalised) or uint8 NHWC RGB -> bf16 [N,H/2,W/2,16] space-to-depth stem input"""
    _require_cuda(x)
    lib = _lib.load()
    if x.dtype == torch.uint8:
        n, h, w, c = x.shape
        fmt = _lib.IN_U8_NHWC
    elif x.dtype == torch.float32:
        n, c, h, w = x.shape
        fmt = _lib.IN_F32_NCHW
    else:
        raise TypeError(f"prep_input: unsupported dtype {x.dtype}")
    if c != 3:
        raise ValueError("prep_input expects 3 channels")
    x = x.contiguous()
    out = torch.empty(n, h // 2, w // 2, 16, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.uwm_prep_input(x.data_ptr(), fmt, n, h, w, out.data_ptr(), _stream()), "uwm_prep_input")
    return out


# ---- training-step glue (csrc/uwm_train.cu) ------------------------------------------------------------------------
_BN_WS: dict = {}


def _bn_workspace(device: torch.device) -> torch.Tensor:
    """UWM_BN_WS_SLOTS (32) x 2 x 2048 fp64 cross-block sums per (device, stream); the kernels leave them zeroed."""
    key = (device.index, _stream())
    ws = _BN_WS.get(key)
    if ws is None:
        ws = _BN_WS[key] = torch.zeros(32 * 2 * 2048, dtype=torch.float64, device=device)
    return ws


def bn_train_forward(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, running_mean: Optional[torch.Tensor],
                     running_var: Optional[torch.Tensor], momentum: float, eps: float, relu: bool,
                     residual: Optional[torch.Tensor] = None):
    """[relu](BatchNorm2d with batch statistics (x) [+ residual]) on dense NHWC bf16; updates the running statistics in
    place.  Returns (y, save) with save = fp32 [4, C]: batch mean, rstd, scale, shift (for bn_train_backward)."""
    _require_cuda(x, gamma, beta, running_mean, running_var, residual)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
    assert residual is None or (residual.is_contiguous() and residual.shape == x.shape and residual.dtype == x.dtype)
    lib = _lib.load()
    c = x.shape[3]
    y = torch.empty_like(x)
    save = torch.empty(4, c, dtype=torch.float32, device=x.device)
    rc = lib.uwm_bn_train_forward_nhwc_bf16(
        x.data_ptr(), x.numel() // c, c, gamma.data_ptr(), beta.data_ptr(),
        running_mean.data_ptr() if running_mean is not None else None,
        running_var.data_ptr() if running_var is not None else None, float(momentum), float(eps),
        residual.data_ptr() if residual is not None else None, int(relu), y.data_ptr(), save.data_ptr(),
        _bn_workspace(x.device).data_ptr(), _stream())
    _lib.check(rc, "uwm_bn_train_forward_nhwc_bf16")
    return y, save


def bn_train_backward(dy: torch.Tensor, x: torch.Tensor, y: Optional[torch.Tensor], save: torch.Tensor, relu: bool,
                      has_residual: bool):
    """Backward of bn_train_forward: (dx, d_residual or None, dgamma, dbeta)."""
    _require_cuda(dy, x, y, save)
    assert dy.is_contiguous() and x.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype == torch.bfloat16
    assert not has_residual or (y is not None and y.is_contiguous())
    lib = _lib.load()
    c = x.shape[3]
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if has_residual else None
    grads = torch.empty(4, c, dtype=torch.float32, device=x.device)        # dgamma, dbeta, 2 x coefficient scratch
    rc = lib.uwm_bn_train_backward_nhwc_bf16(
        dy.data_ptr(), x.data_ptr(), y.data_ptr() if y is not None else None, x.numel() // c, c, save.data_ptr(),
        int(relu), int(has_residual), dx.data_ptr(), dres.data_ptr() if dres is not None else None,
        grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(), _bn_workspace(x.device).data_ptr(), _stream())
    _lib.check(rc, "uwm_bn_train_backward_nhwc_bf16")
    return dx, dres, grads[0], grads[1]


def upsample2x_backward(dy: torch.Tensor) -> torch.Tensor:
    """Backward of nearest 2x: dy [N,2h,2w,C] (may be the leading channel slice of a wider buffer) -> [N,h,w,C]."""
    _require_cuda(dy)
    lib = _lib.load()
    n, h2, w2, c = dy.shape
    dx = torch.empty(n, h2 // 2, w2 // 2, c, dtype=torch.bfloat16, device=dy.device)
    rc = lib.uwm_upsample2x_backward_nhwc_bf16(dy.data_ptr(), n, h2 // 2, w2 // 2, c, _pitch(dy), dx.data_ptr(), c, _stream())
    _lib.check(rc, "uwm_upsample2x_backward_nhwc_bf16")
    return dx


def maxpool3x3s2_backward(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Backward of MaxPool2d(3, 2, 1): x [N,h,w,C] is the pool's input, dy [N,(h-1)//2+1,(w-1)//2+1,C] -> dx like x."""
    _require_cuda(dy, x)
    assert x.is_contiguous() and dy.is_contiguous() and x.dtype == dy.dtype == torch.bfloat16
    n, h, w, c = x.shape
    assert tuple(dy.shape) == (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), (dy.shape, x.shape)
    lib = _lib.load()
    dx = torch.empty_like(x)
    rc = lib.uwm_maxpool3x3s2_backward_nhwc_bf16(dy.data_ptr(), x.data_ptr(), n, h, w, c, dx.data_ptr(), _stream())
    _lib.check(rc, "uwm_maxpool3x3s2_backward_nhwc_bf16")
    return dx


def pack_train_weights(weight: torch.Tensor, with_dgrad: bool):
    """fp32 [Cout,Cin,kh,kw] -> (bf16 [Cout, kh*kw*Cin] forward operand, bf16 [Cin, kh*kw*Cout] flipped / transposed
    data-gradient operand or None), one launch (``training.dgrad_weights`` is the torch restatement)."""
    _require_cuda(weight)
    assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.dim() == 4
    cout, cin, kh, kw = weight.shape
    lib = _lib.load()
    fwd = torch.empty(cout, kh * kw * cin, dtype=torch.bfloat16, device=weight.device)
    dgrad = torch.empty(cin, kh * kw * cout, dtype=torch.bfloat16, device=weight.device) if with_dgrad else None
    rc = lib.uwm_pack_train_weights(weight.data_ptr(), cout, cin, kh, kw, fwd.data_ptr(),
                                    dgrad.data_ptr() if dgrad is not None else None, _stream())
    _lib.check(rc, "uwm_pack_train_weights")
    return fwd, dgrad


def copy_channels(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NHWC bf16 channel slice -> channel slice (either side may be the slice of a wider buffer; ``dst`` None: a new dense
    tensor)."""
    _require_cuda(src, dst)
    n, h, w, c = src.shape
    if dst is None:
        dst = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=src.device)
    assert src.dtype == dst.dtype == torch.bfloat16 and tuple(dst.shape) == (n, h, w, c)
    lib = _lib.load()
    rc = lib.uwm_copy_channels_nhwc_bf16(src.data_ptr(), n * h * w, c, _pitch(src), dst.data_ptr(), _pitch(dst), _stream())
    _lib.check(rc, "uwm_copy_channels_nhwc_bf16")
    return dst
