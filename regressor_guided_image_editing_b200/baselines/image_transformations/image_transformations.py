"""Drop-in for src/baselines/image_transformations/image_transformations.py (reference file:line cited per function).

Same names, argument meaning and return types as the reference; each `apply_*` is a torch.autograd.Function whose
forward/backward launch the hand-written CUDA kernels in librgie.so (csrc/filters.cu).  The reference's own parameter
clamps (e.g. `torch.clamp(saturation_param, min=0)`, :98) stay in torch so autograd sees the same graph.

Batch semantics: like the reference, a 0-d / single parameter set is shared by all images of `im`; additionally a
parameter tensor with a leading dimension equal to the batch is applied per image (the batched engine uses that).
"""
from __future__ import annotations

import math

import torch

from ... import _lib, ops


def _per_image(param: torch.Tensor, n: int, B: int):
    """-> (flat float32 contiguous tensor, stride in floats between images (0 = shared))."""
    p = param.reshape(-1) if param.numel() == n else param.reshape(B, n)
    p = p.contiguous().float()
    return p, (0 if p.dim() == 1 else n)


class _Filter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, param, kind, n):
        if not im.is_cuda:
            raise _lib.RgieError("apply_* needs CUDA tensors: there is no CPU path in this package")
        imc = im.contiguous().float()
        p, stride = _per_image(param.to(imc.device), n, imc.shape[0])
        out = ops.filter_fwd(kind, imc, p, stride)
        ctx.save_for_backward(imc, p)
        ctx.kind, ctx.n, ctx.stride, ctx.pshape = kind, n, stride, param.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        imc, p = ctx.saved_tensors
        B = imc.shape[0]
        gp = torch.empty(B, ctx.n, dtype=torch.float32, device=imc.device)
        gin = ops.filter_bwd(ctx.kind, imc, gout.contiguous().float(), p, ctx.stride, gp, ctx.n)
        gparam = gp.sum(0) if ctx.stride == 0 else gp
        return gin, gparam.reshape(ctx.pshape), None, None


def _as_tensor(param, im):
    if not isinstance(param, torch.Tensor):
        param = torch.as_tensor(param, dtype=torch.float32, device=im.device)
    return param


def apply_exposure(im, exposure_param):                      # :146-153 -> img_trans_torch_diff.py:60-64
    return _Filter.apply(im, _as_tensor(exposure_param, im), _lib.F_EXPOSURE, 1)


def apply_saturation(im, saturation_param):                  # :91-98
    return _Filter.apply(im, torch.clamp(_as_tensor(saturation_param, im), min=0), _lib.F_SATURATION, 1)


def apply_tone_curve_adjustment(im, tone_param):             # :80-88 -> img_trans_torch_diff.py:6-19
    return _Filter.apply(im, tone_param, _lib.F_TONE, 8)


def apply_color_curve_adjustment(im, color_param):           # :69-77
    return _Filter.apply(im, color_param, _lib.F_COLOR, 24)


def apply_contrast(im, contrast_param):                      # :101-109
    return _Filter.apply(im, _as_tensor(contrast_param, im), _lib.F_CONTRAST, 1)


def apply_sharpening(im, sharp_param):                       # :188-195
    return _Filter.apply(im, torch.clamp(_as_tensor(sharp_param, im), min=0), _lib.F_SHARP, 1)


def apply_gaussian_blur(im, blur_param, kernel_size=(25, 25)):   # :112-123
    if tuple(kernel_size) != (25, 25):
        raise _lib.RgieError("only the reference's (25, 25) kernel is implemented")
    return _Filter.apply(im, torch.clamp(_as_tensor(blur_param, im), min=0), _lib.F_BLUR, 1)


def apply_scale(im, scale_param):                            # :209-221
    if scale_param.size(-1) != 4:
        raise _lib.RgieError("apply_scale: expected (sx, sy, cx, cy) per image")
    return _Filter.apply(im, scale_param, _lib.F_SCALE, 4)


def apply_white_balance(im, white_balance_param):            # :126-133 -> img_trans_torch_diff.py:51-57
    return _Filter.apply(im, _as_tensor(white_balance_param, im), _lib.F_WB, 1)


def apply_brightness(im, brightness_param):                  # :136-143
    return _Filter.apply(im, torch.clamp(_as_tensor(brightness_param, im), min=0, max=1), _lib.F_BRIGHT, 1)


def apply_black_white(im, bw_param):                         # :156-163 -> img_trans_torch_diff.py:67-70
    return _Filter.apply(im, _as_tensor(bw_param, im), _lib.F_BW, 1)


def apply_hue(im, hue_param):                                # :166-173
    return _Filter.apply(im, torch.clamp(_as_tensor(hue_param, im), min=-math.pi, max=math.pi), _lib.F_HUE, 1)


def apply_gamma(im, gamma_param):                            # :176-185
    return _Filter.apply(im, torch.clamp(_as_tensor(gamma_param, im), min=0), _lib.F_GAMMA, 1)


def apply_affine_transformation(im, matrices):               # :198-206
    """kornia.geometry.transform.affine(im, matrices, padding_mode='border') + clamp; matrices [2,3] or [B,2,3]."""
    return _Filter.apply(im, matrices, _lib.F_AFFINE, 6)


_DISPATCH = {
    "exposure": apply_exposure, "saturation": apply_saturation, "tone": apply_tone_curve_adjustment,
    "color": apply_color_curve_adjustment, "contrast": apply_contrast, "sharp": apply_sharpening,
    "blur": apply_gaussian_blur, "scale": apply_scale,
    "gamma": apply_gamma, "wb": apply_white_balance, "bright": apply_brightness, "bw": apply_black_white, "hue": apply_hue,
    "affine": apply_affine_transformation,
}


def apply_params(im, params):                                # :7-66
    """Applies parameters to image in dict order; returns the list of stage outputs (last one carries grad)."""
    param_names = list(params.keys())
    im_list = []
    for i, name in enumerate(param_names):
        if name not in _DISPATCH:
            raise _lib.RgieError(f"filter '{name}' is not implemented natively yet (SURVEY.md 8f rank 1)")
        im = _DISPATCH[name](im, params[name])     # the trailing clamp(0,1) of :60 is fused into every kernel
        im_list.append(im if i == len(param_names) - 1 else im.detach().clone())
    return im_list
