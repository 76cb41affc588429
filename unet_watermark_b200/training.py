"""Optional training step of the path (BASELINE.json configs[4]: Unet-resnet34, 512x512, batch 16 per GPU, Dice+BCE,
NCCL gradient all-reduce).  Semantics source: reference src/train.py:68-127 (``train_epoch``: ``model.train()``,
forward, loss, backward, optimizer step), :265-277 (Adam(lr, weight_decay)), src/utils/losses.py, src/configs/config.py.

What runs where
  forward   every convolution whose channel counts are multiples of 16 (all but the 3-channel stem and the 1-channel
            head) runs on the hand-written tcgen05 implicit-GEMM kernel through the C ABI (``uwm_conv2d_nhwc_bf16``:
            bf16 NHWC, fp32 accumulation), as a ``torch.autograd.Function``.  BatchNorm runs in TRAIN mode (batch
            statistics, running-stat update) - so it cannot be folded into the conv as the inference plan does - on the
            hand-written HBM-bound kernels of csrc/uwm_train.cu, fused with the ReLU and the residual add (``_BNFn``;
            parameters and running statistics are the model's own ``nn.BatchNorm2d`` modules'); the decoder's nearest
            upsample is written straight into the concat buffer (``_UpCatFn``); the stem's max-pool runs on the inference
            path's kernel, its backward on ``maxpool3x3s2_bwd_kernel`` (``_MaxPoolFn``).
  backward  data gradients of the stride-1 convs (39 of the 45 convs of Unet-resnet34 that run on the C ABI) run on the same tcgen05
            kernel: dgrad of a 'same' conv is the conv of gy with the flipped, in/out-transposed filter (``dgrad_weights``;
            ``UWM_NATIVE_DGRAD=0`` turns it off).  Weight gradients, and the data gradients of the stride-2 convs, are
            ``aten.convolution_backward`` (cuDNN) on the saved bf16 tensors; a hand-written wgrad (K = pixels, both
            operands MN-major) is the remaining part of SURVEY.md §8 N4.
  launch    zero_grad + forward + loss + backward replay as ONE CUDA graph per input shape (``TrainStep``; ~1000 launches
            and ~200 Python autograd-function calls per step are host-bound otherwise; ``UWM_TRAIN_GRAPH=0``: eager).
  exchange  :class:`GradBuckets` - flat fp32 gradient buckets (parameters own views into them) all-reduced over NCCL:
            after the graph when the step is graph-replayed (98 MB over NVSwitch is well under a millisecond), or, eager,
            as soon as the backward pass has produced every gradient of a bucket, overlapping the rest of the backward;
            the exposed part (the wait after the backward) is measured with CUDA events.  Adam is torch's fused kernel.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops


_ZERO_BIAS: dict = {}


def _zero_bias(c: int, device: torch.device) -> torch.Tensor:
    """fp32 zeros [c] for the bias-free convs (one tensor per (c, device) instead of a fill launch per conv call)."""
    key = (c, device.index)
    z = _ZERO_BIAS.get(key)
    if z is None:
        z = _ZERO_BIAS[key] = torch.zeros(c, dtype=torch.float32, device=device)
    return z


class _ConvFn(torch.autograd.Function):
    """y = conv2d(x, w) (no bias) on the tcgen05 kernel; x, y: NCHW-shaped channels_last bf16 tensors."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding):
        cout, cin, kh, kw = weight.shape
        w = weight.detach()
        wd = None
        if native_pack_enabled() and w.dtype == torch.float32 and w.is_contiguous():
            # one launch: bf16 forward operand + (when the data gradient runs on the tcgen05 kernel) its flipped,
            # in/out-transposed twin (csrc/uwm_train.cu pack_train_weights_kernel)
            wp, wd = ops.pack_train_weights(w, ctx.needs_input_grad[0] and native_dgrad_applies(weight.shape, stride, padding))
            w16 = wp.view(cout, kh, kw, cin).permute(0, 3, 1, 2)  # [cout,cin,kh,kw] with channels_last strides, no copy
        else:
            # the channels_last bf16 filter IS the UWM_PACK_TAPS layout ([cout][kh*kw][cin]) and the layout cuDNN's
            # weight-gradient call wants in the backward
            w16 = w.to(torch.bfloat16, memory_format=torch.channels_last)
            wp = w16.permute(0, 2, 3, 1).reshape(cout, -1)
            if not wp.is_contiguous():                            # (1x1 filters: strides of size-1 dims are ambiguous)
                wp = wp.contiguous()
        xn = x.detach().permute(0, 2, 3, 1)                       # NHWC view of the channels_last tensor
        if not xn.is_contiguous():
            xn = xn.contiguous()
        y = ops.conv2d(xn, wp, _zero_bias(cout, x.device), kh, kw, stride, padding, relu=False)
        ctx.save_for_backward(x, w16, wd)
        ctx.conf = (stride, padding)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x, w16, wd = ctx.saved_tensors
        stride, padding = ctx.conf
        gy = gy.contiguous(memory_format=torch.channels_last)
        need_gx = ctx.needs_input_grad[0]
        gx = None
        if need_gx and native_dgrad_applies(w16.shape, stride, padding):
            # data gradient of a stride-1 'same' conv = the conv of gy with the spatially flipped, in/out-transposed
            # filter: the forward's own tcgen05 kernel, K = taps x Cout, N = Cin
            cin = w16.shape[1]
            gx = ops.conv2d(gy.permute(0, 2, 3, 1), wd if wd is not None else dgrad_weights(w16), _zero_bias(cin, gy.device),
                            w16.shape[2], w16.shape[3], 1, padding, relu=False).permute(0, 3, 1, 2)
            need_gx = False
        g2, gw, _ = torch.ops.aten.convolution_backward(
            gy, x, w16.contiguous(memory_format=torch.channels_last), None, [stride, stride], [padding, padding], [1, 1],
            False, [0, 0], 1, [need_gx, True, False])
        return (gx if gx is not None else g2), gw.float() if gw is not None else None, None, None


def native_pack_enabled() -> bool:
    return os.environ.get("UWM_NATIVE_PACK", "1") != "0"


def native_dgrad_enabled() -> bool:
    return os.environ.get("UWM_NATIVE_DGRAD", "1") != "0"


def native_dgrad_applies(wshape, stride: int, padding: int) -> bool:
    """Data gradients that run on ``uwm_conv2d_nhwc_bf16``: stride-1 'same' convs (every 3x3 / 1x1 conv of the network
    but the three stride-2 stage entries and their 1x1 downsamples, whose transposed conv stays on cuDNN)."""
    cout, cin, kh, kw = wshape
    return (native_dgrad_enabled() and stride == 1 and kh == kw and kh % 2 == 1 and padding == kh // 2
            and cin % 16 == 0 and cout % 16 == 0)


def dgrad_weights(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,kh,kw] -> UWM_PACK_TAPS weights [Cin][kh*kw][Cout] of the conv that maps gy to gx:
    ``gx[n,ci,h,w] = sum_{co,r,s} gy[n,co,h+pad-r,w+pad-s] * w[co,ci,r,s]``, i.e. tap (r',s') = (kh-1-r, kw-1-s) of a
    'same' conv over gy with the roles of Cin and Cout swapped."""
    cin = w.shape[1]
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, -1).contiguous()


def _conv(x: torch.Tensor, m: nn.Conv2d) -> torch.Tensor:
    cout, cin = m.weight.shape[:2]
    if cin % 16 == 0 and cout % 16 == 0 and m.bias is None:
        return _ConvFn.apply(x, m.weight, m.stride[0], m.padding[0])
    # 3-channel stem / 1-channel head: not expressible on the 16-channel-granular single-operator ABI
    return F.conv2d(x, m.weight.to(x.dtype), None if m.bias is None else m.bias.to(x.dtype), m.stride, m.padding)


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """Dense NHWC view of an NCHW-shaped tensor (a copy only when it is not channels_last already)."""
    v = t.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


class _BNFn(torch.autograd.Function):
    """[relu](BatchNorm2d with batch statistics (x) [+ residual]) on csrc/uwm_train.cu: three HBM-bound launches forward
    (per-channel sums, finalise + running statistics, normalise + add + ReLU), three backward (ReLU mask + sums, finalise,
    dx / d residual).  torch's channels_last batch-norm kernels took 9.3 ms of a 26.8 ms step at 512x512 x 16."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, bn, relu):
        xn = _nhwc(x.detach())
        rn = _nhwc(residual.detach()) if residual is not None else None
        momentum = bn.momentum
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:                                   # cumulative moving average (torch semantics)
                momentum = 1.0 / float(bn.num_batches_tracked)
        y, save = ops.bn_train_forward(xn, weight.detach(), bias.detach(),
                                       bn.running_mean if bn.track_running_stats else None,
                                       bn.running_var if bn.track_running_stats else None,
                                       momentum if momentum is not None else 0.0, bn.eps, relu, rn)
        if bn.track_running_stats:                                  # written through raw pointers: tell autograd's version
            for t in (bn.running_mean, bn.running_var):              # counters (Unet._token() watches them)
                torch.autograd.graph.increment_version(t)
        ctx.save_for_backward(xn, y if rn is not None else None, save)
        ctx.relu, ctx.has_res = bool(relu), rn is not None
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, y, save = ctx.saved_tensors
        dx, dres, dgamma, dbeta = ops.bn_train_backward(_nhwc(gy), xn, y, save, ctx.relu, ctx.has_res)
        return (dx.permute(0, 3, 1, 2), dgamma, dbeta, dres.permute(0, 3, 1, 2) if dres is not None else None, None, None)


def native_bn_applies(x: torch.Tensor, bn: nn.BatchNorm2d) -> bool:
    c = x.shape[1]
    return (os.environ.get("UWM_NATIVE_BN", "1") != "0" and bn.training and bn.affine and x.is_cuda
            and x.dtype == torch.bfloat16 and c % 8 == 0 and c <= 2048 and bn.weight.dtype == torch.float32)


def _bn_relu(x, bn: nn.BatchNorm2d, relu: bool = True, residual: Optional[torch.Tensor] = None):
    """relu(bn(x) [+ residual]) with the module's parameters and running statistics (train mode: batch statistics)."""
    if native_bn_applies(x, bn):
        return _BNFn.apply(x, bn.weight, bn.bias, residual, bn, relu)
    y = bn(x)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


class _UpCatFn(torch.autograd.Function):
    """cat([interpolate(x, 2, 'nearest'), skip], 1) (smp DecoderBlock.forward) written once: the upsample kernel stores
    into the leading channels of the concat buffer, the skip is copied behind it; backward = 2x2 sums of the leading
    channels of the gradient + a view of the rest."""

    @staticmethod
    def forward(ctx, x, skip):
        xn = _nhwc(x.detach())
        n, h, w, cx = xn.shape
        cs = skip.shape[1] if skip is not None else 0
        out = torch.empty(n, 2 * h, 2 * w, cx + cs, dtype=xn.dtype, device=xn.device)
        ops.upsample2x(xn, out=out[..., :cx] if cs else out)
        if cs:
            ops.copy_channels(_nhwc(skip.detach()), out[..., cx:])
        ctx.cx, ctx.cs = cx, cs
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        gn = _nhwc(g)
        dx = ops.upsample2x_backward(gn[..., :ctx.cx] if ctx.cs else gn).permute(0, 3, 1, 2)
        # a dense copy of the skip's slice: its consumers (BatchNorm backward, the sum with the encoder-side gradient)
        # would each re-gather the strided view through torch's scalar copy kernel
        dskip = ops.copy_channels(gn[..., ctx.cx:]).permute(0, 3, 1, 2) if ctx.cs else None
        return dx, dskip


def _upcat(x: torch.Tensor, skip: Optional[torch.Tensor]) -> torch.Tensor:
    cs = skip.shape[1] if skip is not None else 0
    if (os.environ.get("UWM_NATIVE_UPCAT", "1") != "0" and x.is_cuda and x.dtype == torch.bfloat16
            and x.shape[1] % 8 == 0 and cs % 8 == 0):
        return _UpCatFn.apply(x, skip)
    y = F.interpolate(x, scale_factor=2, mode="nearest")
    if skip is not None:
        y = torch.cat([y, skip], dim=1)                             # upsampled first, skip second
    return y.contiguous(memory_format=torch.channels_last)


class _MaxPoolFn(torch.autograd.Function):
    """MaxPool2d(3, 2, 1) of the stem (torchvision ResNet.maxpool): forward on the inference path's kernel (bit-exact with
    ``F.max_pool2d``), backward on ``maxpool3x3s2_bwd_kernel`` with torch's first-maximum arg-max rule recomputed from
    the saved input instead of an index tensor."""

    @staticmethod
    def forward(ctx, x):
        xn = _nhwc(x.detach())
        ctx.save_for_backward(xn)
        return ops.maxpool3x3s2(xn).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        (xn,) = ctx.saved_tensors
        return ops.maxpool3x3s2_backward(_nhwc(gy), xn).permute(0, 3, 1, 2)


def _maxpool(x: torch.Tensor) -> torch.Tensor:
    if (os.environ.get("UWM_NATIVE_POOL", "1") != "0" and x.is_cuda and x.dtype == torch.bfloat16
            and x.shape[1] % 8 == 0 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0):
        return _MaxPoolFn.apply(x)
    return F.max_pool2d(x, 3, 2, 1)


def forward_train(model, x: torch.Tensor) -> torch.Tensor:
    """smp ``Unet.forward`` with BatchNorm in the module's current mode and autograd enabled.
    x: fp32 [B,3,H,W] (ImageNet-normalised) -> fp32 logits [B,1,H,W]."""
    if not x.is_cuda:
        raise RuntimeError("unet_watermark_b200 trains on CUDA (sm_100a) only; there is no CPU fallback")
    model.check_input_shape(x.shape[-2], x.shape[-1])
    enc = model.encoder
    y = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = _bn_relu(_conv(y, enc.conv1), enc.bn1)
    feats = [y]
    y = _maxpool(y)
    for layer in (enc.layer1, enc.layer2, enc.layer3, enc.layer4):
        for blk in layer:
            idt = y
            if blk.downsample is not None:
                idt = _bn_relu(_conv(y, blk.downsample[0]), blk.downsample[1], relu=False)
            if hasattr(blk, "conv3"):                               # Bottleneck
                t = _bn_relu(_conv(y, blk.conv1), blk.bn1)
                t = _bn_relu(_conv(t, blk.conv2), blk.bn2)
                y = _bn_relu(_conv(t, blk.conv3), blk.bn3, residual=idt)
            else:                                                   # BasicBlock
                t = _bn_relu(_conv(y, blk.conv1), blk.bn1)
                y = _bn_relu(_conv(t, blk.conv2), blk.bn2, residual=idt)
        feats.append(y)
    skips = feats[::-1]                                             # layer4, layer3, layer2, layer1, stem
    y = skips[0]
    for i, blk in enumerate(model.decoder.blocks):
        y = _upcat(y, skips[i + 1] if i + 1 < len(skips) else None)
        y = _bn_relu(_conv(y, blk.conv1[0]), blk.conv1[1])
        y = _bn_relu(_conv(y, blk.conv2[0]), blk.conv2[1])
    logits = _conv(y, model.segmentation_head[0]).float()
    if model.activation_name == "sigmoid":
        logits = torch.sigmoid(logits)
    return logits


class GradBuckets:
    """Bucketed gradient all-reduce overlapped with the backward pass.

    Parameters are grouped, in reverse registration order (the order the backward produces their gradients), into
    flat fp32 buckets of ~``bucket_mb``; ``p.grad`` is a view into its bucket, so autograd accumulates in place.  A
    post-accumulate hook counts a bucket's finished gradients and launches ``all_reduce(async)`` on the last one.
    ``finish()`` waits for the outstanding collectives and averages."""

    def __init__(self, params, bucket_mb: float = 25.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._hooks = []
        cap = int(bucket_mb * 1024 * 1024 / 4)
        cur: List[torch.nn.Parameter] = []
        n = 0
        groups = []
        for p in reversed(self.params):
            if cur and n + p.numel() > cap:
                groups.append(cur); cur, n = [], 0
            cur.append(p); n += p.numel()
        if cur:
            groups.append(cur)
        for bi, g in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in g), dtype=torch.float32, device=g[0].device)
            off = 0
            for p in g:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self.buckets.append(flat)
        self._sizes = [len(g) for g in groups]
        self._ready = [0] * len(groups)
        self._works = []
        self.overlap = True              # False: hooks only count (a CUDA-graph-captured backward); finish(all_buckets=True)
        self.bytes = sum(b.numel() * 4 for b in self.buckets)

    def zero_grad(self):
        for b in self.buckets:
            b.zero_()
        self._ready = [0] * len(self.buckets)
        self._works = []

    def _on_grad(self, p):
        bi = self._bucket_of[p]
        self._ready[bi] += 1
        if self._ready[bi] == self._sizes[bi] and self.world > 1 and self.overlap:
            self._works.append(dist.all_reduce(self.buckets[bi], group=self.group, async_op=True))

    def finish(self, all_buckets: bool = False):
        """Wait for every bucket's all-reduce and turn the sums into means.  ``all_buckets``: nothing was launched
        from the hooks (graph replay) - reduce every bucket now."""
        if self.world > 1:
            for i, n in enumerate(self._ready):                  # parameters that received no gradient this step
                if all_buckets or n != self._sizes[i]:
                    self._works.append(dist.all_reduce(self.buckets[i], group=self.group, async_op=True))
            for w in self._works:
                w.wait()
            for b in self.buckets:
                b.div_(self.world)
        self._works = []

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


class TrainStep:
    """One optimisation step as the reference runs it (src/train.py:82-107), data-parallel over the process group."""

    def __init__(self, model, cfg=None, lr: Optional[float] = None, weight_decay: Optional[float] = None,
                 criterion: Optional[nn.Module] = None, bucket_mb: float = 25.0, use_graph: Optional[bool] = None):
        from .config import get_cfg_defaults
        from .losses import dice_bce_from_config
        self.cfg = cfg if cfg is not None else get_cfg_defaults()
        self.model = model
        self.criterion = criterion if criterion is not None else dice_bce_from_config(self.cfg)
        # reference src/train.py:265-270: Adam(lr = TRAIN.LR, weight_decay = TRAIN.WEIGHT_DECAY)
        # (fused = the same update as one multi-tensor launch per step instead of ~20 foreach launches)
        on_cuda = all(p.is_cuda for p in model.parameters())
        self.optimizer = torch.optim.Adam(model.parameters(), lr=self.cfg.TRAIN.LR if lr is None else lr,
                                          weight_decay=self.cfg.TRAIN.WEIGHT_DECAY if weight_decay is None else weight_decay,
                                          fused=True if on_cuda else None)
        self.buckets = GradBuckets(model.parameters(), bucket_mb=bucket_mb)
        self.exposed_ms: List[float] = []
        # CUDA graph of zero_grad + forward + loss + backward per (shape, dtype) of the inputs: ~1000 kernel launches and
        # ~200 Python autograd-function calls per step are host-bound once the kernels are fast (UWM_TRAIN_GRAPH=0: eager)
        self.use_graph = os.environ.get("UWM_TRAIN_GRAPH", "1") != "0" if use_graph is None else bool(use_graph)
        if any(isinstance(m, nn.BatchNorm2d) and m.momentum is None for m in model.modules()):
            self.use_graph = False            # cumulative-average BatchNorm reads num_batches_tracked on the host
        self.graph_warmup = 2                 # eager steps on the capture stream before capturing
        self._graphs: dict = {}
        self.replayed_native_launches = 0     # libuwm_b200.so kernels run by graph replays (the library counts captures only)
        self._seen: dict = {}
        self._stream: Optional[torch.cuda.Stream] = None

    def _forward_backward(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        self.buckets.zero_grad()                                   # optimizer.zero_grad(): grads are bucket views
        out = self.model(images)
        loss = self.criterion(out, masks)
        loss.backward()
        return loss.detach()

    def _exchange_and_update(self, time_exchange: bool, all_buckets: bool):
        if time_exchange:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.buckets.finish(all_buckets)
            e1.record()
            self._pending = (e0, e1)
        else:
            self.buckets.finish(all_buckets)
        self.optimizer.step()

    def step(self, images: torch.Tensor, masks: torch.Tensor, time_exchange: bool = False) -> torch.Tensor:
        """images: fp32 [B,3,H,W] normalised; masks: {0,1} [B,H,W] or [B,1,H,W].  Returns the (detached) loss."""
        self.model.train()
        if masks.dim() == 3:
            masks = masks.unsqueeze(1)
        if not (self.use_graph and images.is_cuda):
            self.buckets.overlap = True
            loss = self._forward_backward(images, masks)
            self._exchange_and_update(time_exchange and images.is_cuda, False)
            return loss
        return self._step_graph(images, masks, time_exchange)

    def _step_graph(self, images: torch.Tensor, masks: torch.Tensor, time_exchange: bool) -> torch.Tensor:
        """The same step with zero_grad + forward + loss + backward replayed from one CUDA graph; the gradient exchange
        (NCCL, after the graph instead of from hooks inside the backward: 98 MB over NVSwitch is ~0.5 ms) and Adam follow it."""
        key = (tuple(images.shape), images.dtype, tuple(masks.shape), masks.dtype, images.device.index)
        cur = torch.cuda.current_stream(images.device)
        if self._stream is None:
            self._stream = torch.cuda.Stream(images.device)
        entry = self._graphs.get(key)
        if entry is None:
            n = self._seen.get(key, 0)
            self._seen[key] = n + 1
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                self.buckets.overlap = False
                if n < self.graph_warmup:                            # eager on the capture stream (allocator, cuDNN plans)
                    loss = self._forward_backward(images, masks)
                else:
                    from . import _lib
                    sx, st = images.clone(), masks.clone()
                    g = torch.cuda.CUDAGraph()
                    l0 = _lib.load().uwm_kernel_launch_count()
                    with torch.cuda.graph(g, stream=self._stream):
                        sl = self._forward_backward(sx, st)
                    # launches of libuwm_b200.so inside the graph (the library counted them at capture = this first replay)
                    entry = self._graphs[key] = (g, sx, st, sl, int(_lib.load().uwm_kernel_launch_count() - l0))
                    g.replay()
                    loss = sl.clone()
            cur.wait_stream(self._stream)
            self._exchange_and_update(time_exchange, True)
            return loss
        g, sx, st, sl, n_native = entry
        sx.copy_(images, non_blocking=True)
        st.copy_(masks, non_blocking=True)
        g.replay()
        self.replayed_native_launches += n_native
        loss = sl.clone()
        self._exchange_and_update(time_exchange, True)
        return loss

    def exposed_exchange_ms(self) -> float:
        """Device time between the end of the backward pass and the last averaged bucket of the latest step timed with
        ``time_exchange=True`` (the part of the all-reduce the backward did not hide)."""
        e0, e1 = self._pending
        e1.synchronize()
        return e0.elapsed_time(e1)


class DevicePrefetcher:
    """Host batches -> device batches with the upload of batch i+1 running on its own stream under step i: the reference's
    ``DataLoader(pin_memory=True)`` + ``images.to(device)`` / ``masks.to(device)`` (src/train.py:84-87) with the copy
    taken off the step's stream.  ``batches`` yields ``(images, masks)`` host tensors (pinned, or the copy is
    synchronous); iterating yields device tensors that stay valid until the next ``depth - 1`` items have been taken.
    On a CPU device it passes the batches through."""

    def __init__(self, batches, device, depth: int = 2):
        self.batches = batches
        self.device = torch.device(device)
        self.depth = max(int(depth), 2)
        self._copy: Optional[torch.cuda.Stream] = None
        self._slots: List[Optional[dict]] = [None] * self.depth     # staging tensors, kept across iterations

    def __iter__(self):
        dev = self.device
        if dev.type != "cuda":
            for x, t in self.batches:
                yield x.to(dev), t.to(dev)
            return
        if self._copy is None:
            self._copy = torch.cuda.Stream(dev)
        copy = self._copy
        slots = self._slots

        def issue(k: int, batch):
            x, t = batch
            s = slots[k]
            if s is None or s["x"].shape != x.shape or s["x"].dtype != x.dtype or s["t"].shape != t.shape or s["t"].dtype != t.dtype:
                s = slots[k] = {"x": torch.empty(x.shape, dtype=x.dtype, device=dev),
                                "t": torch.empty(t.shape, dtype=t.dtype, device=dev),
                                "ready": torch.cuda.Event(), "free": None}
                copy.wait_stream(torch.cuda.current_stream(dev))     # the allocation's stream
            with torch.cuda.stream(copy):
                if s["free"] is not None:
                    copy.wait_event(s["free"])                      # the step that read this slot last has finished
                s["x"].copy_(x, non_blocking=True)
                s["t"].copy_(t, non_blocking=True)
                s["ready"].record(copy)

        it = iter(self.batches)
        nxt = next(it, None)
        if nxt is None:
            return
        issue(0, nxt)
        i = 0
        while True:
            k = i % self.depth
            nxt = next(it, None)
            if nxt is not None:
                issue((i + 1) % self.depth, nxt)
            torch.cuda.current_stream(dev).wait_event(slots[k]["ready"])
            try:
                yield slots[k]["x"], slots[k]["t"]
            finally:                                                # also when the consumer stops early
                free = torch.cuda.Event()
                free.record(torch.cuda.current_stream(dev))
                slots[k]["free"] = free
            if nxt is None:
                return
            i += 1
