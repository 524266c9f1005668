"""Drop-in for src/baselines/models/EmotionPredictionModel.py:10-54 (`load_model_eval`).

Returns an nn.Sequential with the reference's stage order: [Resize(aa) -> ReplicateAndCrop(10 random crops, normalise)
-> resnet50.eval()] fused into one native module, then MeanReplicatedCrops, then the activation.  The fused module is a
torch.autograd.Function around librgie.so: forward = antialiased resize + crop packing + tcgen05 (bf16) or CUDA-core
(fp32) resnet50; backward = the input-gradient path only (the reference's `requires_grad=True` also produces weight
gradients that nothing reads, SURVEY.md 8a R3 -- they are not computed here).
"""
from __future__ import annotations

import os
from typing import Dict

import torch
import torch.nn as nn

from ... import _lib, ops
from .utilities.MeanReplicatedCrops import MeanReplicatedCrops
from .utilities.ReplicateAndCrop import draw_crop_offsets

DEFAULT_PRECISION = os.environ.get("RGIE_PRECISION", "bf16")


class _RegressorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, mod):
        if not img.is_cuda:
            raise _lib.RgieError("the native regressor needs CUDA tensors: there is no CPU path in this package")
        x = img.contiguous().float()
        B, _, H, W = x.shape
        if mod.input_size is not None:
            oh, ow = ops.resize_output_size(H, W, mod.input_size)
        else:
            oh, ow = H, W
        rs = mod._resize(H, W, oh, ow)
        xr = rs.fwd(x)
        reps = mod.num_replications
        offs = draw_crop_offsets(B, oh, ow, mod.crop_size, reps).to(x.device)
        reg = mod._regressor(B * reps)
        logits = reg.forward(xr, offs, normalize=mod.normalize)
        ctx.rs, ctx.reg, ctx.shape = rs, reg, (B, 3, oh, ow)
        # the native handle keeps the activations (sign bits, pool argmax) of its LAST forward only: remember which
        # forward this was, and what it ran on, so that backward can restore the state if another forward came between
        # (ValenceArousalLoss.forward(fake, real_imgs=...) runs the model twice before backpropagating the first call)
        ctx.generation, ctx.fwd_args, ctx.normalize = reg.generation, (xr, offs), mod.normalize
        mod.last_offsets = offs
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=dlogits.device)
        if ctx.reg.generation != ctx.generation:
            xr, offs = ctx.fwd_args
            ctx.reg.forward(xr, offs, normalize=ctx.normalize)        # same inputs, same offsets: identical activations
        ctx.reg.backward(dlogits.contiguous().float(), dx)
        return ctx.rs.bwd(dx), None


class NativeCropResNet50(nn.Module):
    """Resize(input_size, antialias) -> ReplicateAndCrop(crop_size, normalize, 10) -> resnet50 (eval): [B,3,H,W] ->
    [B*10, num_classes] logits."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], input_size, crop_size, normalize, num_replications=10,
                 precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.input_size, self.crop_size, self.normalize = input_size, crop_size, normalize
        self.num_replications, self.precision = num_replications, precision
        self._sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self._folded = None
        self._regs, self._resizes = {}, {}
        self.last_offsets = None

    def _regressor(self, n_crops: int) -> "ops.Regressor":
        if n_crops not in self._regs:
            if self._folded is None:
                self._folded = ops.fold_resnet50(self._sd)
            self._regs = {}       # one workspace at a time
            self._regs[n_crops] = ops.Regressor(self._sd, n_crops, self.crop_size, self.precision, folded=self._folded)
        return self._regs[n_crops]

    def _resize(self, h, w, oh, ow) -> "ops.Resize":
        key = (h, w, oh, ow)
        if key not in self._resizes:
            self._resizes[key] = ops.Resize(h, w, oh, ow)
        return self._resizes[key]

    def forward(self, x):
        return _RegressorFn.apply(x, self)


def load_model_eval(path_to_model, num_classes, input_size=None, crop_size=None, normalize=False,
                    activation_function=None, is_ten_crop=False, requires_grad=False, precision: str = DEFAULT_PRECISION):
    """Same signature as the reference (:10-11) plus `precision` ('bf16' tcgen05 | 'fp32' parity mode).
    `path_to_model` is a torch.save'd torchvision resnet50 state_dict with fc -> num_classes (or the dict itself)."""
    sd = path_to_model if isinstance(path_to_model, dict) else torch.load(path_to_model, map_location="cpu")
    if sd["fc.weight"].shape[0] != num_classes:
        raise _lib.RgieError(f"state_dict has {sd['fc.weight'].shape[0]} outputs, expected {num_classes}")
    if not is_ten_crop or crop_size is None:
        raise _lib.RgieError("only the ten-crop configuration used by ValenceArousalLoss is implemented natively")
    modules = [NativeCropResNet50(sd, input_size, crop_size, normalize, 10, precision), MeanReplicatedCrops(10)]
    if activation_function is not None:
        modules.append(activation_function)
    return nn.Sequential(*modules)
