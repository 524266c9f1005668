"""Drop-in for src/guidance_classifier/MiduClassifier.py: same constructor, hook and `forward(latents, t, prompt_embeds)`.

`self.model` stays the reference's nn.Sequential (state_dict compatible, :121-161; SD and SDXL variants) but its
arithmetic runs in the native head (csrc/midu.cu): a torch.autograd.Function maps the hooked mid-block feature
[B,1280,8,8] (SD) / [B,1280,32,32] (SDXL) to [B,n_out] and returns
d/d(feature) on backward, so `torch.autograd.grad(loss, latents)` in the caller's sampling loop
(pipelines/InversionResamplingStableDiffusionPipeline.py:132-134) continues into the caller's own UNet.
The UNet itself is third-party (diffusers) and out of scope here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib
from .._lib import check, ptr, stream_ptr
from .GuidanceClassifier import GuidanceClassifier

DEFAULT_PRECISION = os.environ.get("RGIE_PRECISION", "bf16")


class _MiduHandle:
    """Owns one native head handle; destroyed when the last reference (the cache or an autograd ctx) goes away, so a
    backward can never run on a handle that a later forward with another batch size / new weights replaced."""

    def __init__(self, h: C.c_void_p, n_out: int):
        self.h, self.n_out, self.generation = h, n_out, 0

    def __del__(self):
        try:
            if self.h:
                _lib.load().rgie_midu_destroy(self.h)
                self.h = C.c_void_p(0)
        except Exception:
            pass


class NativeMiduHead:
    """librgie.so handle for the SD / SDXL head, rebuilt when the batch size or the weights change."""

    def __init__(self, model: nn.Sequential, precision: str, is_sdxl: bool = False):
        self.model, self.precision, self.is_sdxl = model, precision, is_sdxl
        self._cur, self._key = None, None

    def _weights_version(self):
        return tuple(p._version for p in self.model.parameters()) + tuple(p.data_ptr() for p in self.model.parameters())

    def handle(self, batch: int, hw: int):
        key = (batch, hw, self._weights_version())
        if key != self._key:
            self._cur = None                       # an autograd ctx that still needs the old handle keeps it alive
            sd = self.model.state_dict()
            layers = ["0", "3", "6", "9", "13", "15"] if self.is_sdxl else ["0", "3", "7", "9"]   # convs then the two linears
            names = [f"{l}.{k}" for l in layers for k in ("weight", "bias")]
            arrs = [np.ascontiguousarray(sd[n].detach().float().cpu().numpy()) for n in names]
            pt = (C.c_void_p * len(arrs))(*[a.ctypes.data_as(C.c_void_p) for a in arrs])
            n_out = int(sd[layers[-1] + ".weight"].shape[0])
            h = C.c_void_p(0)
            check(_lib.load().rgie_midu_create(pt, len(arrs), n_out, batch, hw, _lib.PRECISIONS[self.precision],
                                               C.byref(h)), "rgie_midu_create")
            self._cur, self._key, self.n_out = _MiduHandle(h, n_out), key, n_out
        return self._cur

    def close(self):
        self._cur, self._key = None, None


class _MiduHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, head):
        if not feat.is_cuda:
            raise _lib.RgieError("the native MiDU head needs CUDA tensors: there is no CPU path in this package")
        f = feat.contiguous().float()
        B, Cc, H, W = f.shape
        if Cc != 1280 or H != W:
            raise _lib.RgieError(f"unexpected mid-block feature shape {tuple(f.shape)}")
        hd = head.handle(B, H)
        pred = torch.empty(B, hd.n_out, dtype=torch.float32, device=f.device)
        check(_lib.load().rgie_midu_forward(hd.h, ptr(f), B, ptr(pred), stream_ptr(f.device)), "rgie_midu_forward")
        hd.generation += 1
        # the handle holds the activations of its last forward only: keep the object (not the raw pointer) and the input
        ctx.hd, ctx.generation, ctx.feat = hd, hd.generation, f
        ctx.shape, ctx.in_dtype = f.shape, feat.dtype
        return pred

    @staticmethod
    def backward(ctx, dpred):
        hd = ctx.hd
        if hd.generation != ctx.generation:        # another forward ran on this handle in between: restore its state
            scratch = torch.empty(ctx.shape[0], hd.n_out, dtype=torch.float32, device=dpred.device)
            check(_lib.load().rgie_midu_forward(hd.h, ptr(ctx.feat), ctx.shape[0], ptr(scratch), stream_ptr(dpred.device)),
                  "rgie_midu_forward")
            hd.generation += 1
            ctx.generation = hd.generation
        dfeat = torch.empty(ctx.shape, dtype=torch.float32, device=dpred.device)
        check(_lib.load().rgie_midu_backward(hd.h, ptr(dpred.contiguous().float()), ptr(dfeat),
                                             stream_ptr(dpred.device)), "rgie_midu_backward")
        return dfeat.to(ctx.in_dtype), None


class MiduClassifier(GuidanceClassifier):
    def __init__(self, pipe, device: str, ckp_path: str = None, num_outputs: int = 1, is_minimized: bool = True,
                 is_sdxl: bool = False, precision: str = DEFAULT_PRECISION):
        super().__init__(device)
        self.pipe = pipe
        self.is_minimized = is_minimized
        self.pipe.unet.mid_block.register_forward_hook(self.__hook_fn)
        self.model = self._create_midu_classifier(self.device, num_outputs, is_sdxl)
        if ckp_path is not None:
            self.model.load_state_dict(torch.load(ckp_path))
            self.model.eval()
        self.criterion = nn.MSELoss()
        self.reference_value = None
        self._native = NativeMiduHead(self.model, precision, is_sdxl)

    def head(self, feature: Tensor) -> Tensor:
        """self.model(feature) evaluated by the native kernels (differentiable w.r.t. feature)."""
        return _MiduHeadFn.apply(feature, self._native)

    def forward(self, latents: Tensor, t: float, prompt_embeds: Tensor = None) -> Tensor:              # :37-50
        self._set_midu_layer(latents, t, prompt_embeds)
        return self._calculate_score(self.pipe.unet.mid_block.output.to(torch.float32), self.head,
                                     self.device, self.is_minimized, self.reference_value)

    def get_loss(self, latents, labels, t, prompts):                                                   # :52-64
        self._set_midu_layer_no_grad(latents, t, prompts)
        outputs = self.head(self.pipe.unet.mid_block.output.to(torch.float32))
        return self.criterion(outputs, labels), outputs

    def predict_score(self, latents, t, prompts):                                                      # :66-78
        self._set_midu_layer_no_grad(latents, t, prompts)
        with torch.no_grad():
            return self.head(self.pipe.unet.mid_block.output.to(torch.float32))

    def _set_midu_layer_no_grad(self, latents, t, prompts):                                            # :80-95
        # prompt -> embedding conversion lives in the reference's pipelines/diff_utils.py (diffusers glue, out of scope):
        # callers pass ready-made embeddings here
        with torch.no_grad():
            self._set_midu_layer(latents, t, prompts)

    def _set_midu_layer(self, latents, t, prompt_embeds):                                              # :97-115
        latents = self.pipe.scheduler.scale_model_input(latents, t)
        if isinstance(prompt_embeds, (list, tuple)) and len(prompt_embeds) == 2 and isinstance(prompt_embeds[1], dict):
            latents = latents.to(prompt_embeds[0].dtype)
            self.pipe.unet(latents, t, encoder_hidden_states=prompt_embeds[0], cross_attention_kwargs=None,
                           added_cond_kwargs=prompt_embeds[1])
        else:
            self.pipe.unet(latents, t, encoder_hidden_states=prompt_embeds)

    @staticmethod
    def __hook_fn(module, input, output):
        module.output = output

    @staticmethod
    def _create_midu_classifier(device, num_outputs=10, is_sdxl=False):                                # :121-161
        if is_sdxl:                                                                                    # :125-143
            m = nn.Sequential(
                nn.Conv2d(1280, 512, kernel_size=3, padding=1), nn.ReLU(), nn.MaxPool2d(2, 2),
                nn.Conv2d(512, 256, kernel_size=3, padding=1), nn.ReLU(), nn.MaxPool2d(2, 2),
                nn.Conv2d(256, 128, kernel_size=3, padding=1), nn.ReLU(), nn.MaxPool2d(2, 2),
                nn.Conv2d(128, 64, kernel_size=3, padding=1), nn.ReLU(), nn.MaxPool2d(2, 2),
                nn.Flatten(), nn.Linear(64 * 2 * 2, 128), nn.ReLU(), nn.Linear(128, num_outputs))
            return m.to(device)
        m = nn.Sequential(
            nn.Conv2d(1280, 256, kernel_size=3, padding=1), nn.ReLU(), nn.MaxPool2d(2, 2),
            nn.Conv2d(256, 128, kernel_size=3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d(output_size=(2, 2)),
            nn.Flatten(), nn.Linear(128 * 4, 64), nn.ReLU(), nn.Linear(64, num_outputs))
        return m.to(device)

    @staticmethod
    def _calculate_score(x, m, device, is_minimized=True, reference_value=None):
        return Tensor(0)
