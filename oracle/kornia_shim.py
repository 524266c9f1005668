"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Pure-torch restatement of the nine kornia 0.8.2 entry points the reference's filter chain calls
(/root/reference/src/baselines/image_transformations/image_transformations.py:98,109,122,143,173,185,195,205,221).

kornia (pinned 0.8.2 in /root/reference/uv.lock:588-590) is NOT vendored in the reference tree and is NOT installed
in this image (no network), so this module restates kornia's published algorithms from the library's documented
behaviour.  PARITY UNPINNED for these nine functions: there is no kornia build here to diff against.  Everything
else in the oracle is validated against the reference's own Python (see oracle/ref_harness.py).

`install()` registers the shim as `kornia` in sys.modules so that the reference's unmodified `apply_params`
runs on top of it.
"""
from __future__ import annotations

import math
import sys
import types

import torch
import torch.nn.functional as F
from torch import Tensor


# ----------------------------------------------------------------------------------------------------------------
# kornia.color
# ----------------------------------------------------------------------------------------------------------------
def rgb_to_hsv(image: Tensor, eps: float = 1e-8) -> Tensor:
    """kornia.color.rgb_to_hsv: h in [0, 2pi], s, v in [0, 1]."""
    max_rgb, argmax_rgb = image.max(-3)
    min_rgb, _ = image.min(-3)
    deltac = max_rgb - min_rgb

    v = max_rgb
    s = deltac / (max_rgb + eps)

    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = torch.unbind((max_rgb.unsqueeze(-3) - image), dim=-3)

    h1 = bc - gc
    h2 = (rc - bc) + 2.0 * deltac
    h3 = (gc - rc) + 4.0 * deltac

    h = torch.stack((h1, h2, h3), dim=-3) / deltac.unsqueeze(-3)
    h = torch.gather(h, dim=-3, index=argmax_rgb.unsqueeze(-3)).squeeze(-3)
    h = (h / 6.0) % 1.0
    h = 2.0 * math.pi * h
    return torch.stack((h, s, v), dim=-3)


def hsv_to_rgb(image: Tensor) -> Tensor:
    """kornia.color.hsv_to_rgb."""
    h = image[..., 0, :, :] / (2 * math.pi)
    s = image[..., 1, :, :]
    v = image[..., 2, :, :]

    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    one = torch.tensor(1.0, device=image.device, dtype=image.dtype)
    p = v * (one - s)
    q = v * (one - f * s)
    t = v * (one - (one - f) * s)

    hi = hi.long()
    indices = torch.stack([hi, hi + 6, hi + 12], dim=-3)
    out = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    out = torch.gather(out, -3, indices)
    return out


def rgb_to_grayscale(image: Tensor) -> Tensor:
    """kornia.color.rgb_to_grayscale with the default (0.299, 0.587, 0.114) weights."""
    r = image[..., 0:1, :, :]
    g = image[..., 1:2, :, :]
    b = image[..., 2:3, :, :]
    return 0.299 * r + 0.587 * g + 0.114 * b


# ----------------------------------------------------------------------------------------------------------------
# kornia.enhance
# ----------------------------------------------------------------------------------------------------------------
def _as_factor(factor, image: Tensor) -> Tensor:
    if not isinstance(factor, Tensor):
        factor = torch.as_tensor(factor, device=image.device, dtype=image.dtype)
    return factor.to(image.device, image.dtype)


def adjust_saturation(image: Tensor, factor) -> Tensor:
    factor = _as_factor(factor, image)
    for _ in image.shape[1:]:
        factor = torch.unsqueeze(factor, dim=-1)
    x_hsv = rgb_to_hsv(image)
    h, s, v = torch.chunk(x_hsv, chunks=3, dim=-3)
    s_out = torch.clamp(s * factor, min=0, max=1)
    return hsv_to_rgb(torch.cat([h, s_out, v], dim=-3))


def adjust_hue(image: Tensor, factor) -> Tensor:
    factor = _as_factor(factor, image)
    for _ in image.shape[1:]:
        factor = torch.unsqueeze(factor, dim=-1)
    x_hsv = rgb_to_hsv(image)
    h, s, v = torch.chunk(x_hsv, chunks=3, dim=-3)
    divisor = 2 * math.pi
    h_out = torch.fmod(h + factor, divisor)
    return hsv_to_rgb(torch.cat([h_out, s, v], dim=-3))


def adjust_gamma(input: Tensor, gamma, gain=1.0) -> Tensor:
    gamma = _as_factor(gamma, input)
    gain = _as_factor(gain, input)
    for _ in range(len(input.shape) - len(gamma.shape)):
        gamma = torch.unsqueeze(gamma, dim=-1)
    for _ in range(len(input.shape) - len(gain.shape)):
        gain = torch.unsqueeze(gain, dim=-1)
    x_adjust = gain * torch.pow(input, gamma)
    return torch.clamp(x_adjust, 0.0, 1.0)


def adjust_brightness(image: Tensor, factor, clip_output: bool = True) -> Tensor:
    factor = _as_factor(factor, image)
    while len(factor.shape) != len(image.shape):
        factor = factor[..., None]
    img_adjust = image + factor
    if clip_output:
        img_adjust = img_adjust.clamp(min=0.0, max=1.0)
    return img_adjust


def adjust_contrast_with_mean_subtraction(image: Tensor, factor) -> Tensor:
    factor = _as_factor(factor, image)
    while len(factor.shape) != len(image.shape):
        factor = factor[..., None]
    if not bool((factor >= 0).any()):
        raise ValueError("Contrast factor must be positive.")
    if image.shape[-3] == 3:
        img_mean = rgb_to_grayscale(image).mean((-2, -1), True)
    else:
        img_mean = image.mean()
    img_adjust = image * factor + img_mean * (1 - factor)
    return img_adjust.clamp(min=0.0, max=1.0)


def _blend_one(input1: Tensor, input2: Tensor, factor: Tensor) -> Tensor:
    if factor == 0.0:
        return input1
    if factor == 1.0:
        return input2
    diff = (input2 - input1) * factor
    res = input1 + diff
    if factor > 0.0 and factor < 1.0:
        return res
    return torch.clamp(res, 0, 1)


def sharpness(input: Tensor, factor) -> Tensor:
    factor = _as_factor(factor, input)
    if len(factor.size()) != 0 and factor.shape != torch.Size([input.size(0)]):
        raise AssertionError("factor must be 0-d or of shape (B,)")
    kernel = (
        torch.as_tensor([[1, 1, 1], [1, 5, 1], [1, 1, 1]], dtype=input.dtype, device=input.device)
        .view(1, 1, 3, 3)
        .repeat(input.size(1), 1, 1, 1)
        / 13
    )
    degenerate = F.conv2d(input, kernel, bias=None, stride=1, groups=input.size(1))
    degenerate = torch.clamp(degenerate, 0.0, 1.0)
    mask = torch.ones_like(degenerate)
    padded_mask = F.pad(mask, [1, 1, 1, 1])
    padded_degenerate = F.pad(degenerate, [1, 1, 1, 1])
    result = torch.where(padded_mask == 1, padded_degenerate, input)
    if len(factor.size()) == 0:
        return _blend_one(result, input, factor)
    return torch.stack([_blend_one(result[i], input[i], factor[i]) for i in range(len(factor))])


# ----------------------------------------------------------------------------------------------------------------
# kornia.filters
# ----------------------------------------------------------------------------------------------------------------
def _gaussian1d(window_size: int, sigma: Tensor) -> Tensor:
    batch_size = sigma.shape[0]
    x = (torch.arange(window_size, device=sigma.device, dtype=sigma.dtype) - window_size // 2).expand(batch_size, -1)
    if window_size % 2 == 0:
        x = x + 0.5
    gauss = torch.exp(-x.pow(2.0) / (2 * sigma.pow(2.0)))
    return gauss / gauss.sum(-1, keepdim=True)


def _filter2d(input: Tensor, kernel: Tensor, border_type: str) -> Tensor:
    # kernel: (B or 1, kH, kW), depthwise, cross-correlation, 'same' output via explicit padding
    b, c, h, w = input.shape
    tmp_kernel = kernel[:, None, ...].to(device=input.device, dtype=input.dtype)
    tmp_kernel = tmp_kernel.expand(-1, c, -1, -1)
    height, width = tmp_kernel.shape[-2:]
    pad = [width // 2, width - 1 - width // 2, height // 2, height - 1 - height // 2]
    input = F.pad(input, pad, mode=border_type)
    tmp_kernel = tmp_kernel.reshape(-1, 1, height, width)
    input = input.view(-1, tmp_kernel.size(0), input.size(-2), input.size(-1))
    output = F.conv2d(input, tmp_kernel, groups=tmp_kernel.size(0), padding=0, stride=1)
    return output.view(b, c, h, w)


def gaussian_blur2d(input: Tensor, kernel_size, sigma, border_type: str = "reflect", separable: bool = True) -> Tensor:
    if isinstance(sigma, tuple):
        sigma = torch.tensor([sigma], device=input.device, dtype=input.dtype)
    else:
        sigma = sigma.to(device=input.device, dtype=input.dtype)
    ky, kx = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    bs = sigma.shape[0]
    kernel_x = _gaussian1d(kx, sigma[:, 1].view(bs, 1))
    kernel_y = _gaussian1d(ky, sigma[:, 0].view(bs, 1))
    out_x = _filter2d(input, kernel_x[..., None, :], border_type)
    return _filter2d(out_x, kernel_y[..., None], border_type)


# ----------------------------------------------------------------------------------------------------------------
# kornia.geometry.transform
# ----------------------------------------------------------------------------------------------------------------
def _normal_transform_pixel(height: int, width: int, device, dtype, eps: float = 1e-14) -> Tensor:
    tr_mat = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], device=device, dtype=dtype)
    width_denom = eps if width == 1 else width - 1.0
    height_denom = eps if height == 1 else height - 1.0
    tr_mat[0, 0] = tr_mat[0, 0] * 2.0 / width_denom
    tr_mat[1, 1] = tr_mat[1, 1] * 2.0 / height_denom
    return tr_mat.unsqueeze(0)


def _normalize_homography(dst_pix_trans_src_pix: Tensor, dsize_src, dsize_dst) -> Tensor:
    src_h, src_w = dsize_src
    dst_h, dst_w = dsize_dst
    device, dtype = dst_pix_trans_src_pix.device, dst_pix_trans_src_pix.dtype
    src_norm_trans_src_pix = _normal_transform_pixel(src_h, src_w, device, dtype)
    src_pix_trans_src_norm = torch.linalg.inv(src_norm_trans_src_pix)
    dst_norm_trans_dst_pix = _normal_transform_pixel(dst_h, dst_w, device, dtype)
    return dst_norm_trans_dst_pix @ (dst_pix_trans_src_pix @ src_pix_trans_src_norm)


def warp_affine(src: Tensor, M: Tensor, dsize, mode: str = "bilinear", padding_mode: str = "zeros",
                align_corners: bool = True) -> Tensor:
    B, C, H, W = src.size()
    M_3x3 = F.pad(M, [0, 0, 0, 1], "constant", value=0.0)
    M_3x3[..., -1, -1] += 1.0
    dst_norm_trans_src_norm = _normalize_homography(M_3x3, (H, W), dsize)
    src_norm_trans_dst_norm = torch.linalg.inv(dst_norm_trans_src_norm)
    grid = F.affine_grid(src_norm_trans_dst_norm[:, :2, :], [B, C, dsize[0], dsize[1]], align_corners=align_corners)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


def affine(tensor: Tensor, matrix: Tensor, mode: str = "bilinear", padding_mode: str = "zeros",
           align_corners: bool = True) -> Tensor:
    is_unbatched = tensor.ndimension() == 3
    if is_unbatched:
        tensor = torch.unsqueeze(tensor, dim=0)
    matrix = matrix.expand(tensor.shape[0], -1, -1)
    height, width = tensor.shape[-2], tensor.shape[-1]
    warped = warp_affine(tensor, matrix, (height, width), mode, padding_mode, align_corners)
    if is_unbatched:
        warped = torch.squeeze(warped, dim=0)
    return warped


def get_rotation_matrix2d(center: Tensor, angle: Tensor, scale: Tensor) -> Tensor:
    ang_rad = angle * math.pi / 180.0
    cos_a, sin_a = torch.cos(ang_rad), torch.sin(ang_rad)
    rotation_matrix = torch.stack([cos_a, sin_a, -sin_a, cos_a], dim=-1).view(*angle.shape, 2, 2)
    scaling_matrix = torch.zeros((2, 2), device=center.device, dtype=center.dtype).fill_diagonal_(1).repeat(
        rotation_matrix.size(0), 1, 1)
    scaling_matrix = scaling_matrix * scale.unsqueeze(dim=2).repeat(1, 1, 2)
    scaled_rotation = rotation_matrix @ scaling_matrix
    alpha = scaled_rotation[:, 0, 0]
    beta = scaled_rotation[:, 0, 1]
    x = center[..., 0]
    y = center[..., 1]
    batch_size = center.shape[0]
    one = torch.tensor(1.0, device=center.device, dtype=center.dtype)
    M = torch.zeros(batch_size, 2, 3, device=center.device, dtype=center.dtype)
    M[..., 0:2, 0:2] = scaled_rotation
    M[..., 0, 2] = (one - alpha) * x - beta * y
    M[..., 1, 2] = beta * x + (one - alpha) * y
    return M


def scale(tensor: Tensor, scale_factor: Tensor, center: Tensor | None = None, mode: str = "bilinear",
          padding_mode: str = "zeros", align_corners: bool = True) -> Tensor:
    if len(scale_factor.shape) == 1:
        scale_factor = scale_factor.repeat(1, 2)
    if center is None:
        h, w = tensor.shape[-2:]
        center = torch.tensor([(w - 1) / 2.0, (h - 1) / 2.0], device=tensor.device, dtype=tensor.dtype)
    center = center.expand(tensor.shape[0], -1)
    scale_factor = scale_factor.expand(tensor.shape[0], 2)
    angle = torch.zeros(scale_factor.shape[:1], device=scale_factor.device, dtype=scale_factor.dtype)
    scaling_matrix = get_rotation_matrix2d(center, angle, scale_factor)
    return affine(tensor, scaling_matrix[..., :2, :3], mode, padding_mode, align_corners)


# ----------------------------------------------------------------------------------------------------------------
def build_module() -> types.ModuleType:
    kornia = types.ModuleType("kornia")
    kornia.__version__ = "0.8.2-shim"
    color = types.ModuleType("kornia.color")
    color.rgb_to_hsv, color.hsv_to_rgb, color.rgb_to_grayscale = rgb_to_hsv, hsv_to_rgb, rgb_to_grayscale
    enhance = types.ModuleType("kornia.enhance")
    for fn in (adjust_saturation, adjust_hue, adjust_gamma, adjust_brightness,
               adjust_contrast_with_mean_subtraction, sharpness):
        setattr(enhance, fn.__name__, fn)
    filters = types.ModuleType("kornia.filters")
    filters.gaussian_blur2d = gaussian_blur2d
    geometry = types.ModuleType("kornia.geometry")
    transform = types.ModuleType("kornia.geometry.transform")
    transform.affine, transform.scale, transform.warp_affine = affine, scale, warp_affine
    transform.get_rotation_matrix2d = get_rotation_matrix2d
    geometry.transform = transform
    kornia.color, kornia.enhance, kornia.filters, kornia.geometry = color, enhance, filters, geometry
    return kornia


def install() -> types.ModuleType:
    """Register the shim under the name `kornia` (only if no real kornia is importable)."""
    try:  # pragma: no cover - real kornia is absent in this image
        import kornia as real  # noqa: F401
        if not getattr(real, "__version__", "").endswith("-shim"):
            return real
    except Exception:
        pass
    k = build_module()
    sys.modules["kornia"] = k
    sys.modules["kornia.color"] = k.color
    sys.modules["kornia.enhance"] = k.enhance
    sys.modules["kornia.filters"] = k.filters
    sys.modules["kornia.geometry"] = k.geometry
    sys.modules["kornia.geometry.transform"] = k.geometry.transform
    return k
