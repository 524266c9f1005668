"""Drop-in for the hot-path functions of src/baselines/optimize_image.py: `optimization` (:56-97) and
`get_condition_from_alpha` (:119-123).

`optimization` keeps the reference signature and semantics (Adam betas (0.9, 0.999), lr ramp :69-75, strict-`<`
best-x tracking BEFORE the update :78-81, returns best_x :97).  Two execution paths, both on the device:
  * fused: when the objective is this package's `objective_function_parametric` over the default 8-filter list with a
    native ValenceArousalLoss and no reconstruction / discriminator term, the whole loop runs inside
    engine.ParametricEditEngine (one CUDA graph per step, no host sync); crop draws are consumed from torch's global
    CPU generator in the reference's order.
  * generic: any other objective callable is evaluated through autograd each step; the Adam update and the best-x
    snapshot are the fused rgie_adam_step kernel (loss compared on the device, so the reference's two per-step host
    syncs disappear).

`compute_clip_loss` (:151-183) is the reconstruction term of the script's default objective (weight_recon = 1.0).  Its
image tower is OpenAI's `clip` package (third party, like diffusers' UNet: SURVEY.md 8f rank 3) -- used when importable, or
any object with `.encode_image` assigned to `CLIP_MODEL`; everything around it (resize to 224, range mapping, feature
normalisation, cosine of the first pair) is here, so `weight_recon > 0` runs through the generic path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, ops
from ..engine import ParametricEditEngine, lr_schedule


def get_condition_from_alpha(alpha, clf, img):                                            # :119-123
    condition = clf.predict_loss_metric(img)
    condition = condition + torch.ones(condition.shape).to(condition.device) * alpha
    return torch.clamp(condition, min=0.0, max=1.0)


CLIP_MODEL = None          # as in the reference (:151): loaded on first use; a caller may assign any module with .encode_image


def _clip_image_tower(device):
    global CLIP_MODEL
    if CLIP_MODEL is None:
        try:
            import clip
        except ImportError as e:
            raise _lib.RgieError("weight_recon > 0 needs the CLIP image tower: install OpenAI's `clip` package (the reference "
                                 "calls clip.load('ViT-B/32'), optimize_image.py:171-172) or assign a model with "
                                 ".encode_image to baselines.optimize_image.CLIP_MODEL; or run with weight_recon=0") from e
        CLIP_MODEL, _ = clip.load("ViT-B/32", device=device)
    return CLIP_MODEL


def _unit(features):
    return features / features.norm(dim=-1, keepdim=True)


def compute_clip_loss(image1, image2):                                                    # :152-183
    """1 - cosine similarity of the CLIP image embeddings of the FIRST image pair.  Both images are resized to 224 x 224
    (torchvision's tensor Resize: antialiased bilinear); if image1 lives in [0, 1] (decided by image1 alone, as in the
    reference) both are mapped to [-1, 1] first."""
    tower = _clip_image_tower(image1.device)
    to_unit_range = bool(image1.min() >= 0)

    def embed(image):
        x = torch.nn.functional.interpolate(image, size=[224, 224], mode="bilinear", align_corners=False, antialias=True)
        if to_unit_range:
            x = (x - 0.5) / 0.5
        return _unit(tower.encode_image(x))

    return 1 - (embed(image1) * embed(image2)).sum(dim=-1)[0]


def _fused_eligible(x0, params, objective_function):
    from .. import optimize_image_param as oip
    if objective_function is not oip.objective_function_parametric:
        return False
    clf, image = params.get("clf"), params.get("image")
    if clf is None or image is None or not image.is_cuda or image.shape[0] != 1:
        return False
    if params.get("weight_recon", 0) > 0 or (params.get("dis") is not None and params.get("weight_dis", 0) > 0):
        return False
    if list(params.get("params", {}).keys()) != oip.DEFAULT_TRANS or x0.numel() != 41:
        return False
    core = clf.model[0] if hasattr(clf, "model") and hasattr(clf.model, "__getitem__") else None
    from .models.EmotionPredictionModel import NativeCropResNet50
    if not isinstance(core, NativeCropResNet50) or clf.output_ixs != [0, 1] or len(clf.model) != 3:
        return False
    if not isinstance(clf.model[2], torch.nn.Sigmoid) or params.get("target") is None:
        return False
    return core.input_size is not None and core.normalize


def _optimization_fused(x0, params, learning_rate, num_steps):
    from .models.utilities.ReplicateAndCrop import draw_crop_offsets
    clf, image = params["clf"], params["image"]
    core = clf.model[0]
    _, _, H, W = image.shape
    # the engine (packed weights + workspace) is cached ON the model object, so it can never outlive it or be handed to
    # another model that happens to get the same id()
    key = (H, W, num_steps, core.precision)
    cached = getattr(core, "_fused_engine", None)
    if cached is None or cached[0] != key:
        core._fused_engine = None
        eng = ParametricEditEngine(core._sd, 1, H, W, num_steps, precision=core.precision,
                                   input_size=core.input_size, crop_size=core.crop_size, folded=core._folded)
        core._fused_engine = (key, eng)
    eng = core._fused_engine[1]
    offs = torch.stack([draw_crop_offsets(1, eng.Hr, eng.Wr, core.crop_size, 10) for _ in range(num_steps)])
    eng.load_problem(image.float().contiguous(), offs.to(image.device), alpha=None, target=params["target"],
                     learning_rate=learning_rate, weight_clf=params["weight_clf"], clf_weight=clf.weight,
                     x0=x0.detach().float().cpu())
    eng.advance(num_steps)
    out = eng.results()
    clf.fake_loss_metric = out["preds"][-1:, 0, :2]
    optimization.last_run = out
    return out["best_x"][0].clone()


def optimization(x0, params, objective_function, learning_rate=0.1, lr_rampdown_length=0.25,
                 lr_rampup_length=0.05, num_steps=100, verbose=False):                      # :56-97
    if not x0.is_cuda:
        raise _lib.RgieError("optimization() needs CUDA tensors: there is no CPU path in this package")
    if (lr_rampdown_length, lr_rampup_length) == (0.25, 0.05) and _fused_eligible(x0, params, objective_function):
        return _optimization_fused(x0, params, learning_rate, num_steps)

    x_opt = x0.clone().detach().float().contiguous().requires_grad_(True)
    m, v = torch.zeros_like(x_opt), torch.zeros_like(x_opt)
    best_x = x_opt.clone().detach()
    best_loss = torch.full((1,), float("inf"), device=x_opt.device)
    best_step = torch.zeros(1, dtype=torch.int32, device=x_opt.device)
    for step in range(num_steps):
        t = step / num_steps
        lr_ramp = min(1.0, (1.0 - t) / lr_rampdown_length)
        lr_ramp = 0.5 - 0.5 * np.cos(lr_ramp * np.pi)
        lr_ramp = lr_ramp * min(1.0, t / lr_rampup_length)
        lr = float(learning_rate * lr_ramp)
        loss = objective_function(x_opt, **params)
        g, = torch.autograd.grad(loss, x_opt)
        # ONE problem whatever the shape of x: the reference compares one scalar loss and snapshots the whole tensor
        # (:78-81), also when x is a batch of style codes [B, 8, 1, 1] -- hence the flat views (adam_step treats a 2-D+
        # tensor as one problem per row, with one loss per row)
        ops.adam_step(x_opt.data.view(-1), g.contiguous().view(-1), m.view(-1), v.view(-1), lr, step + 1,
                      loss=loss.detach().reshape(1).float().contiguous(), best_loss=best_loss, best_x=best_x.view(-1),
                      best_step=best_step, step=step)
        if verbose:
            print(f'[ step {step + 1:>4d}/{num_steps}] [ loss: {float(loss):<5.4f}] [ lr: {float(lr):<5.4f}] ')
    return best_x
