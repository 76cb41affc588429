"""Host-side owner of one ``uwm_model`` plan (C ABI handle) for a fixed network input size.

Folds BatchNorm + packs weights from a state-dict-compatible module, uploads them, and runs
``uwm_model_forward`` on torch-owned device buffers and the current torch stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, packing
from .ops import logit


class Engine:
    def __init__(self, encoder_name: str, decoder_channels, height: int, width: int, max_batch: int,
                 device: torch.device):
        if not torch.cuda.is_available():
            raise RuntimeError("unet_watermark_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        enc = {"resnet34": 34, "resnet50": 50}.get(encoder_name)
        if enc is None:
            raise NotImplementedError(f"B200 path implements resnet34/resnet50 encoders, not {encoder_name!r}")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.h, self.w, self.max_batch = int(height), int(width), int(max_batch)
        self.handle = C.c_void_p()
        dec = (C.c_int * 5)(*[int(c) for c in decoder_channels])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.uwm_model_create(enc, dec, self.h, self.w, self.max_batch, C.byref(self.handle)),
                       "uwm_model_create")
        self.layers: List[_lib.LayerDesc] = []
        for i in range(self.lib.uwm_model_num_layers(self.handle)):
            d = _lib.LayerDesc()
            _lib.check(self.lib.uwm_model_layer_desc(self.handle, i, C.byref(d)))
            self.layers.append(d)
        self.weights_loaded = False

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.uwm_model_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # -- info -------------------------------------------------------------------------------
    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.uwm_model_workspace_bytes(self.handle))

    @property
    def kernels_per_forward(self) -> int:
        return int(self.lib.uwm_model_num_kernels(self.handle))

    @property
    def flops_per_image(self) -> float:
        return float(self.lib.uwm_model_flops_per_image(self.handle))

    # -- weights ----------------------------------------------------------------------------
    def load_weights(self, state: Dict[str, torch.Tensor]):
        """state: smp-layout state dict (``encoder.*``, ``decoder.blocks.*``, ``segmentation_head.0.*``)."""
        for i, d in enumerate(self.layers):
            ck, bk = d.conv_key.decode(), d.bn_key.decode()
            w = state[ck + ".weight"].detach().float().cpu()
            if bk:
                w, b = packing.fold_bn(w, state[bk + ".weight"].cpu(), state[bk + ".bias"].cpu(),
                                       state[bk + ".running_mean"].cpu(), state[bk + ".running_var"].cpu())
            else:
                b = state[ck + ".bias"].detach().float().cpu()
            if d.pack == _lib.PACK_STEM_S2D:
                wp = packing.pack_stem_s2d(w, d.cout_pad)
                bp = packing.pad_bias(b, d.cout_pad)
            elif d.pack == _lib.PACK_TAPS_SKIP_PART:
                wp = packing.pack_taps(w[:, w.shape[1] - d.cin_skip:], d.cout_pad)      # skip half of a two-launch conv1
                bp = packing.pad_bias(b, d.cout_pad)
            elif d.pack == _lib.PACK_UP2X_SHUFFLE_X_PART:
                wp = packing.pack_up2x_shuffle(w[:, :w.shape[1] - d.cin_skip])        # upsampled half, bias already added
                bp = torch.zeros(d.cout_pad, dtype=torch.float32)
            elif d.pack == _lib.PACK_S2_PLANES:
                wp = packing.pack_s2_planes(w)                              # stride-2 conv over parity planes
                bp = packing.pad_bias(b, d.cout_pad)
            elif d.pack == _lib.PACK_S2D_CONV:
                wp = packing.pack_s2d_conv3x3(w, d.cout_pad)               # conv on a space-to-depth tensor
                bp = packing.pad_bias(b, d.cout_pad) if d.cout == 1 else b.detach().float().repeat(4).contiguous()
            elif d.pack == _lib.PACK_UPCAT_SUBPIXEL:
                wp = packing.pack_upcat_subpixel(w, d.cin - d.cin_skip)   # sub-pixel conv1 with a skip source
                bp = b.detach().float().repeat(4).contiguous()
            elif d.pack == _lib.PACK_UP2X_SHUFFLE:
                wp = packing.pack_up2x_shuffle(w)              # [4*cout, 9*cin]: one output group per parity
                bp = b.detach().float().repeat(4).contiguous()
            else:
                wp = packing.pack_taps(w, d.cout_pad)
                bp = packing.pad_bias(b, d.cout_pad)
            assert wp.numel() == d.w_elems and bp.numel() == d.b_elems, (ck, wp.shape, d.w_elems)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.uwm_model_set_layer(self.handle, i, wp.data_ptr(), wp.numel(), bp.data_ptr(),
                                                        bp.numel()), f"uwm_model_set_layer({ck})")
        self.weights_loaded = True

    # -- forward ----------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, want_logits: bool = True, threshold: Optional[float] = None,
                sigmoid_threshold: bool = True, apply_sigmoid: bool = False, use_graph: bool = True,
                logits_out: Optional[torch.Tensor] = None, mask_out: Optional[torch.Tensor] = None,
                sub_batch: int = 0) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """x: fp32 [B,3,H,W] (normalised) or uint8 [B,H,W,3] on this engine's device.
        Returns (fp32 [B,1,H,W] logits or None, uint8 [B,H,W] 0/255 mask or None).

        ``sub_batch`` (0 = off): run the batch as consecutive forwards of at most that many images on slices of the same
        input / output buffers.  Images are independent in eval mode, so the result is bit-identical;
        what changes is that a layer's hand-over tensor (B x 18.9 MB for resnet50 layer1 at 768x768) fits the 126 MB L2
        between its producer and its consumer."""
        if not self.weights_loaded:
            raise RuntimeError("Engine.forward before load_weights")
        if not x.is_cuda:
            raise RuntimeError("Engine.forward needs a CUDA tensor (no CPU fallback)")
        if x.dtype == torch.uint8:
            b, h, w, c = x.shape
            fmt = _lib.IN_U8_NHWC
        elif x.dtype == torch.float32:
            b, c, h, w = x.shape
            fmt = _lib.IN_F32_NCHW
        else:
            raise TypeError(f"unsupported input dtype {x.dtype} (float32 NCHW or uint8 NHWC)")
        if c != 3 or (h, w) != (self.h, self.w):
            raise ValueError(f"engine built for 3x{self.h}x{self.w}, got {c}x{h}x{w}")
        x = x.contiguous()
        logits = mask = None
        if want_logits:
            logits = logits_out if logits_out is not None else torch.empty(b, 1, h, w, dtype=torch.float32, device=x.device)
        thr_logit = 0.0
        if threshold is not None:
            mask = mask_out if mask_out is not None else torch.empty(b, h, w, dtype=torch.uint8, device=x.device)
            thr_logit = logit(threshold) if sigmoid_threshold else float(threshold)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        step = b if not sub_batch or sub_batch >= b else int(sub_batch)
        px, pl, pm = x.data_ptr(), (logits.data_ptr() if logits is not None else 0), (mask.data_ptr() if mask is not None else 0)
        in_img = h * w * 3 * x.element_size()                       # bytes per image: input / fp32 logits / uint8 mask
        for i in range(0, b, step):
            rc = self.lib.uwm_model_forward(self.handle, px + i * in_img, fmt, min(step, b - i),
                                            pl + i * h * w * 4 if pl else None, int(apply_sigmoid),
                                            pm + i * h * w if pm else None, thr_logit, int(use_graph), stream)
            _lib.check(rc, "uwm_model_forward")
        return logits, mask

    def read_tensor(self, name: str, batch: int) -> torch.Tensor:
        """bf16 NHWC copy of a named plan tensor of the last forward (needs UWM_KEEP_ALL=1 at create
        for tensors whose buffer is recycled)."""
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.uwm_model_read_tensor(self.handle, name.encode(), batch, None, 0, C.byref(h), C.byref(w),
                                                  C.byref(c), None))
        out = torch.empty(batch, h.value, w.value, c.value, dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.uwm_model_read_tensor(self.handle, name.encode(), batch, out.data_ptr(),
                                                  out.numel() * 2, C.byref(h), C.byref(w), C.byref(c),
                                                  torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def profile(self, x: torch.Tensor, threshold: float = 0.5):
        """Per-kernel device times of one eager forward: list of (name, ms, flops, bytes)."""
        b = x.shape[0]
        fmt = _lib.IN_U8_NHWC if x.dtype == torch.uint8 else _lib.IN_F32_NCHW
        n = self.kernels_per_forward
        names = C.create_string_buffer(n * 64)
        ms = (C.c_float * n)()
        fl = (C.c_double * n)()
        by = (C.c_double * n)()
        mask = torch.empty(b, self.h, self.w, dtype=torch.uint8, device=x.device)
        rc = self.lib.uwm_model_profile(self.handle, x.contiguous().data_ptr(), fmt, b, None, mask.data_ptr(),
                                        logit(threshold), names, ms, fl, by, n,
                                        torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc, "uwm_model_profile")
        out = []
        for i in range(rc):
            nm = names.raw[i * 64:(i + 1) * 64].split(b"\0", 1)[0].decode()
            out.append((nm, float(ms[i]), float(fl[i]), float(by[i])))
        return out
