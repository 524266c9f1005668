"""Tensor-level wrappers over the C ABI (include/rgie.h).  PyTorch is used only as the container for device memory
and streams; all arithmetic happens in librgie.so.  Every wrapper raises if the extension or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

RESNET50_LAYERS = (3, 4, 6, 3)


def _require_cuda(t: torch.Tensor, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise _lib.RgieError(f"{name} must live on a CUDA device: this package has no CPU path")
    if t.dtype != torch.float32 and t.dtype != torch.int32:
        raise _lib.RgieError(f"{name} must be float32/int32, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.RgieError(f"{name} must be contiguous")


# ----------------------------------------------------------------------------------------------------------------
# filters
# ----------------------------------------------------------------------------------------------------------------
_ws_cache: Dict[tuple, torch.Tensor] = {}


def filter_workspace(B: int, H: int, W: int, device) -> torch.Tensor:
    key = (B, H, W, str(device))
    ws = _ws_cache.get(key)
    if ws is None:
        n = _lib.load().rgie_filter_ws_floats(B, H, W)
        ws = torch.empty(n, dtype=torch.float32, device=device)
        _ws_cache[key] = ws
    return ws


def filter_fwd(kind: int, x: torch.Tensor, p: torch.Tensor, p_stride: int, out: Optional[torch.Tensor] = None,
               ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, "image"); _require_cuda(p, "params")
    B, Cc, H, W = x.shape
    assert Cc == 3
    out = torch.empty_like(x) if out is None else out
    ws = filter_workspace(B, H, W, x.device) if ws is None else ws
    check(_lib.load().rgie_filter_fwd(kind, ptr(x), ptr(out), ptr(p), p_stride, B, H, W, ptr(ws), stream_ptr(x.device)),
          "rgie_filter_fwd")
    return out


def filter_prefix_fwd(x: torch.Tensor, p: torch.Tensor, p_stride: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """exposure -> saturation -> tone -> colour in one pass; p = 34 consecutive effective parameters per image."""
    _require_cuda(x, "image"); _require_cuda(p, "params")
    B, _, H, W = x.shape
    out = torch.empty_like(x) if out is None else out
    check(_lib.load().rgie_filter_prefix_fwd(ptr(x), ptr(out), ptr(p), p_stride, B, H, W, stream_ptr(x.device)),
          "rgie_filter_prefix_fwd")
    return out


def filter_prefix_bwd(x: torch.Tensor, gout: torch.Tensor, p: torch.Tensor, p_stride: int, gp: torch.Tensor, gp_stride: int,
                      ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The 34 parameter gradients of the fused head of the chain (no d(image): its input is the fixed original)."""
    _require_cuda(x, "image"); _require_cuda(gout, "grad"); _require_cuda(p, "params"); _require_cuda(gp, "gparams")
    B, _, H, W = x.shape
    ws = filter_workspace(B, H, W, x.device) if ws is None else ws
    check(_lib.load().rgie_filter_prefix_bwd(ptr(x), ptr(gout), ptr(p), p_stride, ptr(gp), gp_stride, B, H, W, ptr(ws),
                                             stream_ptr(x.device)), "rgie_filter_prefix_bwd")
    return gp


def filter_bwd(kind: int, x: torch.Tensor, gout: torch.Tensor, p: torch.Tensor, p_stride: int, gp: torch.Tensor,
               gp_stride: int, gin: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, "image"); _require_cuda(gout, "grad"); _require_cuda(p, "params"); _require_cuda(gp, "gparams")
    B, _, H, W = x.shape
    gin = torch.empty_like(x) if gin is None else gin
    ws = filter_workspace(B, H, W, x.device) if ws is None else ws
    check(_lib.load().rgie_filter_bwd(kind, ptr(x), ptr(gout), ptr(gin), ptr(p), p_stride, ptr(gp), gp_stride, B, H, W,
                                      ptr(ws), stream_ptr(x.device)), "rgie_filter_bwd")
    return gin


# ----------------------------------------------------------------------------------------------------------------
# resize
# ----------------------------------------------------------------------------------------------------------------
def resize_output_size(h: int, w: int, size: int):
    """torchvision.transforms.functional.resize with an int size: smaller edge -> size."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


class Resize:
    """Antialiased bilinear resize handle (tap tables live on the device)."""

    def __init__(self, in_h: int, in_w: int, out_h: int, out_w: int):
        self.in_h, self.in_w, self.out_h, self.out_w = in_h, in_w, out_h, out_w
        self.identity = (in_h, in_w) == (out_h, out_w)
        self._h = C.c_void_p(0)
        if not self.identity:
            check(_lib.load().rgie_resize_create(in_h, in_w, out_h, out_w, C.byref(self._h)), "rgie_resize_create")
        self._tmp: Dict[tuple, torch.Tensor] = {}

    def __del__(self):
        try:
            if self._h:
                _lib.load().rgie_resize_destroy(self._h)
                self._h = C.c_void_p(0)
        except Exception:
            pass

    def _tmpbuf(self, planes: int, device) -> torch.Tensor:
        key = (planes, str(device))
        if key not in self._tmp:
            self._tmp[key] = torch.empty(planes * self.in_h * self.out_w, dtype=torch.float32, device=device)
        return self._tmp[key]

    def fwd(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.identity:
            return x
        _require_cuda(x, "image")
        B, Cc = x.shape[:2]
        out = torch.empty(B, Cc, self.out_h, self.out_w, dtype=torch.float32, device=x.device) if out is None else out
        check(_lib.load().rgie_resize_fwd(self._h, ptr(x), ptr(out), B * Cc, ptr(self._tmpbuf(B * Cc, x.device)),
                                          stream_ptr(x.device)), "rgie_resize_fwd")
        return out

    def bwd(self, gout: torch.Tensor, gin: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.identity:
            return gout
        _require_cuda(gout, "grad")
        B, Cc = gout.shape[:2]
        gin = torch.empty(B, Cc, self.in_h, self.in_w, dtype=torch.float32, device=gout.device) if gin is None else gin
        check(_lib.load().rgie_resize_bwd(self._h, ptr(gout), ptr(gin), B * Cc, ptr(self._tmpbuf(B * Cc, gout.device)),
                                          stream_ptr(gout.device)), "rgie_resize_bwd")
        return gin


# ----------------------------------------------------------------------------------------------------------------
# regressor
# ----------------------------------------------------------------------------------------------------------------
def fold_resnet50(sd: Dict[str, torch.Tensor], eps: float = 1e-5):
    """Fold eval-mode BatchNorm into the conv weights of a torchvision resnet50 state_dict (float64 -> float32).
    Returns the host fp32 arrays in the order rgie_regressor_create expects (include/rgie.h)."""
    def fold(conv: str, bn: str):
        w = sd[conv + ".weight"].detach().double().cpu()
        g, b = sd[bn + ".weight"].detach().double().cpu(), sd[bn + ".bias"].detach().double().cpu()
        mu, var = sd[bn + ".running_mean"].detach().double().cpu(), sd[bn + ".running_var"].detach().double().cpu()
        s = g / torch.sqrt(var + eps)
        return [np.ascontiguousarray((w * s[:, None, None, None]).float().numpy()),
                np.ascontiguousarray((b - mu * s).float().numpy())]

    arrs = fold("conv1", "bn1")
    for li, nb in enumerate(RESNET50_LAYERS, start=1):
        for bi in range(nb):
            pre = f"layer{li}.{bi}"
            arrs += fold(pre + ".conv1", pre + ".bn1") + fold(pre + ".conv2", pre + ".bn2") + \
                fold(pre + ".conv3", pre + ".bn3")
            if bi == 0:
                arrs += fold(pre + ".downsample.0", pre + ".downsample.1")
    arrs.append(np.ascontiguousarray(sd["fc.weight"].detach().float().cpu().numpy()))
    arrs.append(np.ascontiguousarray(sd["fc.bias"].detach().float().cpu().numpy()))
    return arrs


class Regressor:
    """Native resnet50 valence/arousal regressor on `max_crops` crops (forward + input-gradient backward)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], max_crops: int, crop_size: int = 448,
                 precision: str = "bf16", device=None, folded=None):
        if not torch.cuda.is_available():
            raise _lib.RgieError("no CUDA device: the regressor has no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_classes = int(state_dict["fc.weight"].shape[0])
        self.crop_size, self.max_crops, self.precision = crop_size, max_crops, precision
        arrs = fold_resnet50(state_dict) if folded is None else folded
        self._arrs = arrs
        pt = (C.c_void_p * len(arrs))(*[a.ctypes.data_as(C.c_void_p) for a in arrs])
        self._h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            check(_lib.load().rgie_regressor_create(pt, len(arrs), self.num_classes, crop_size, max_crops,
                                                    _lib.PRECISIONS[precision], C.byref(self._h)),
                  "rgie_regressor_create")
        self._arrs = None
        self.generation = 0          # number of forwards run on this handle (backward belongs to the last one)

    def __del__(self):
        try:
            if self._h:
                _lib.load().rgie_regressor_destroy(self._h)
                self._h = C.c_void_p(0)
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        return int(_lib.load().rgie_regressor_workspace_bytes(self._h))

    def forward(self, img: torch.Tensor, offsets: torch.Tensor, normalize: bool = True,
                logits: Optional[torch.Tensor] = None, step_ptr: Optional[torch.Tensor] = None,
                off_step_stride: int = 0, reps: Optional[int] = None) -> torch.Tensor:
        _require_cuda(img, "image"); _require_cuda(offsets, "offsets")
        B, _, Hr, Wr = img.shape
        reps = int(offsets.shape[-2]) if reps is None else reps
        if logits is None:
            logits = torch.empty(B * reps, self.num_classes, dtype=torch.float32, device=img.device)
        check(_lib.load().rgie_regressor_forward_ex(self._h, ptr(img), B, Hr, Wr, ptr(offsets), ptr(step_ptr),
                                                    off_step_stride, reps, int(normalize), ptr(logits),
                                                    stream_ptr(img.device)), "rgie_regressor_forward")
        # the handle keeps the RAW pointers of the image, the crop offsets and the step counter for backward()
        # (crop_grad_gather_kernel reads them): hold the tensors so that a caller's temporaries cannot be recycled by the
        # caching allocator between forward and backward
        self._fwd_refs = (img, offsets, step_ptr)
        self.generation += 1
        return logits

    def set_input_transform(self, pre_scale: float, pre_shift: float, mean, std) -> None:
        """`normalize=2` of forward(): t = clamp(v * pre_scale + pre_shift, 0, 1); (t - mean_c) / std_c (EmoNet pipeline)."""
        m = (C.c_float * 3)(*[float(v) for v in mean]); s = (C.c_float * 3)(*[float(v) for v in std])
        check(_lib.load().rgie_regressor_set_input_transform(self._h, float(pre_scale), float(pre_shift), m, s),
              "rgie_regressor_set_input_transform")

    def backward(self, dlogits: torch.Tensor, dimg: torch.Tensor) -> torch.Tensor:
        _require_cuda(dlogits, "dlogits"); _require_cuda(dimg, "dimg")
        check(_lib.load().rgie_regressor_backward(self._h, ptr(dlogits), ptr(dimg), stream_ptr(dimg.device)),
              "rgie_regressor_backward")
        return dimg

    def tap(self, name: str, shape: Sequence[int]) -> torch.Tensor:
        out = torch.empty(*shape, dtype=torch.float32, device=self.device)
        n = C.c_long(0)
        check(_lib.load().rgie_regressor_tap(self._h, name.encode(), ptr(out), out.numel(), C.byref(n),
                                             stream_ptr(self.device)), "rgie_regressor_tap")
        assert n.value == out.numel(), (name, n.value, out.numel())
        return out


# ----------------------------------------------------------------------------------------------------------------
# loss head / update
# ----------------------------------------------------------------------------------------------------------------
def va_head(logits: torch.Tensor, B: int, reps: int, sigmoid: bool, target: Optional[torch.Tensor], tv_default: float,
            ta_default: float, use_mask: int, scale: float, preds: torch.Tensor, loss: Optional[torch.Tensor],
            dlogits: Optional[torch.Tensor]) -> None:
    nc = logits.shape[-1]
    check(_lib.load().rgie_va_head(ptr(logits), B, reps, nc, int(sigmoid), ptr(target), tv_default, ta_default, use_mask,
                                   scale, ptr(preds), ptr(loss), ptr(dlogits), stream_ptr(logits.device)), "rgie_va_head")


def adam_scalars(lr: float, k: int, beta1: float = 0.9, beta2: float = 0.999):
    """Host-side float64 scalars of torch's _single_tensor_adam for 1-based step k."""
    bc1 = 1 - beta1 ** k
    bc2 = 1 - beta2 ** k
    return lr / bc1, bc2 ** 0.5


def adam_step(x: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, k: int,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, loss: Optional[torch.Tensor] = None,
              best_loss: Optional[torch.Tensor] = None, best_x: Optional[torch.Tensor] = None,
              best_step: Optional[torch.Tensor] = None, step: int = 0) -> None:
    _require_cuda(x, "x"); _require_cuda(g, "g")
    x2 = x.view(1, -1) if x.dim() == 1 else x.view(x.shape[0], -1)
    B, n = x2.shape
    for name, t in (("loss", loss), ("best_loss", best_loss), ("best_step", best_step)):
        if t is not None and t.numel() != B:
            raise _lib.RgieError(f"adam_step: {name} holds {t.numel()} values for {B} problems (x is [{B}, {n}]: one problem "
                                 f"per row; pass flat views for a single problem)")
    step_size, bc2_sqrt = adam_scalars(lr, k, beta1, beta2)
    check(_lib.load().rgie_adam_step(ptr(x), ptr(g), ptr(m), ptr(v), B, n, step_size, bc2_sqrt, 1.0 - beta1, beta2,
                                     1.0 - beta2, eps, ptr(loss), ptr(best_loss), ptr(best_x), ptr(best_step), step,
                                     stream_ptr(x.device)), "rgie_adam_step")


def guidance_update(x: torch.Tensor, g: torch.Tensor, scale: float, normalize: bool = True,
                    per_problem: Optional[int] = None) -> torch.Tensor:
    """In place: x -= scale * g / (||g|| + 1e-10), the norm taken per problem (default: the whole tensor)."""
    _require_cuda(x, "latents"); _require_cuda(g, "grad")
    per = x.numel() if per_problem is None else per_problem
    n_prob = x.numel() // per
    ws = torch.empty(2 * n_prob * 128, dtype=torch.float32, device=x.device)
    check(_lib.load().rgie_guidance_update(ptr(x), ptr(g), n_prob, per, scale, int(normalize), ptr(ws),
                                           stream_ptr(x.device)), "rgie_guidance_update")
    return x
