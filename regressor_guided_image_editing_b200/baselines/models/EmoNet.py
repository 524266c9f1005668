"""Drop-in for src/baselines/models/EmoNet.py (`load_model_eval`, :33-44): the EmoNet valence regressor -- torchvision
resnet50 with fc -> 1 -- behind its deterministic ten-crop pipeline (:63-88, :110-130):

    [Normalize(.5,.5) if normalize] -> Resize(256, antialias) -> denorm ((x+1)/2, clamp(0,1), *255)
    -> ten crops of 224 (4 corners + centre at (17,17), + their horizontal flips) -> /255 -> ImageNet normalisation
    -> resnet50 -> mean over the 10 crops -> [valence, 0]   (fake arousal column, :91-95)

Native mapping: the resize is the antialiased resample kernel, crops / flips / clamp / normalisation happen while the
conv1 operand is packed (librgie.so `normalize` mode 2, mirrored crops by bit 30 of the `left` offset), resnet50 forward
and input-gradient backward are the tcgen05 (bf16) or CUDA-core (fp32) kernels.  Resize is linear with weights summing
to 1, so Normalize(.5,.5) followed by denorm's (x+1)/2 cancels: the native path resizes the [0,1] image directly (the
two orders differ by fp32 round-off only).
"""
from __future__ import annotations

import math
import os
from typing import Dict

import torch
import torch.nn as nn

from ... import _lib, ops

DEFAULT_PRECISION = os.environ.get("RGIE_PRECISION", "bf16")
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]          # EmoNet.mean / .std (:19-25)
FLIP = 1 << 30


def tencrop_offsets(batch: int, im_size: int = 256, cropped_size: int = 224) -> torch.Tensor:
    """int32 [batch, 10, 2] (top, left | FLIP) in the reference's crop order (EmoNet.py:110-130)."""
    hi = im_size - cropped_size
    center = int(math.floor(hi / 2) + 1)
    base = [(0, 0), (0, hi), (hi, 0), (hi, hi), (center, center)]
    one = [[t, l] for t, l in base] + [[t, l | FLIP] for t, l in base]
    return torch.tensor(one, dtype=torch.int32)[None].repeat(batch, 1, 1).contiguous()


def convert_checkpoint(parameters: Dict) -> Dict[str, torch.Tensor]:
    """EmoNet checkpoint ({'state_dict': {'<prefix>.model.*', 'model.last_linear.*'}}, :50-54) -> torchvision resnet50
    keys with fc -> 1.  A plain torchvision state_dict passes through."""
    sd = parameters["state_dict"] if "state_dict" in parameters else parameters
    if any(k.startswith("conv1.") for k in sd):
        return dict(sd)
    out = {}
    for k, v in sd.items():
        k = ".".join(k.split(".")[1:])                   # drop the DataParallel / wrapper prefix (:51)
        k = k.replace("model.last_linear.", "model.fc.")
        out[k[len("model."):] if k.startswith("model.") else k] = v
    return out


class _EmoNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, mod):
        if not img.is_cuda:
            raise _lib.RgieError("the native regressor needs CUDA tensors: there is no CPU path in this package")
        x = img.contiguous().float()
        B, _, H, W = x.shape
        oh, ow = ops.resize_output_size(H, W, 256)
        rs = mod._resize(H, W, oh, ow)
        xr = rs.fwd(x)
        offs = tencrop_offsets(B).to(x.device)
        reg = mod._regressor(B * 10)
        logits = reg.forward(xr, offs, normalize=2)
        # backward re-reads the crop offsets and (for the clamp mask) the resized image: keep both alive
        ctx.rs, ctx.reg, ctx.shape, ctx.xr, ctx.offs = rs, reg, (B, 3, oh, ow), xr, offs
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=dlogits.device)
        ctx.reg.backward(dlogits.contiguous().float(), dx)
        return ctx.rs.bwd(dx), None


class NativeEmoNet(nn.Module):
    """[B,3,H,W] -> [B*10, 1] logits of the ten crops."""

    def __init__(self, state_dict, normalize: bool, precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.normalize, self.precision = normalize, precision
        self._sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self._folded, self._regs, self._resizes = None, {}, {}

    def _regressor(self, n_crops: int):
        if n_crops not in self._regs:
            if self._folded is None:
                self._folded = ops.fold_resnet50(self._sd)
            self._regs = {n_crops: ops.Regressor(self._sd, n_crops, 224, self.precision, folded=self._folded)}
            # normalize=True: input in [0,1] (Normalize(.5,.5) and denorm cancel); False: input in [-1,1] -> (x+1)/2
            ps, pb = (1.0, 0.0) if self.normalize else (0.5, 0.5)
            self._regs[n_crops].set_input_transform(ps, pb, MEAN, STD)
        return self._regs[n_crops]

    def _resize(self, h, w, oh, ow):
        key = (h, w, oh, ow)
        if key not in self._resizes:
            self._resizes[key] = ops.Resize(h, w, oh, ow)
        return self._resizes[key]

    def forward(self, x):
        return _EmoNetFn.apply(x, self)


class TenCropOutput(nn.Module):
    """tencrop_output_transform_emonet (:91-95): mean over the 10 crops + a zero "arousal" column."""

    def forward(self, output):
        output = output.view(-1, 10).mean(1)
        return torch.stack((output, torch.zeros(output.shape[0]).to(output.device)), dim=1)


def load_model_eval(path_to_model, normalize=False, requires_grad=False, precision: str = DEFAULT_PRECISION):   # :33-44
    params = path_to_model if isinstance(path_to_model, dict) else torch.load(path_to_model, map_location="cpu")
    sd = convert_checkpoint(params)
    if sd["fc.weight"].shape[0] != 1:
        raise _lib.RgieError("EmoNet: expected a resnet50 with a single output (last_linear -> 1)")
    return nn.Sequential(NativeEmoNet(sd, normalize, precision), TenCropOutput())
