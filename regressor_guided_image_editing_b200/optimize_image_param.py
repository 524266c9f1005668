"""Drop-in for the hot-path functions of src/optimize_image_param.py: `init_params` (:121-209),
`initialize_parametric` (:212-234), `objective_function_parametric` (:237-259), `get_params_from_vector` (:262-292),
and the caller right after the loop, `output_transform` (:295-312: evaluation + full-resolution re-render of the edit,
SURVEY.md 8f rank 4).  Script glue (main, dataset loop, JPEG save, discriminator loading) is out of scope (section 8); the CLIP term runs when its
third-party image tower is available (baselines/optimize_image.py::compute_clip_loss).
"""
from __future__ import annotations

import torch

from . import _lib
from .baselines.image_transformations.image_transformations import apply_params

STATS = {}                 # :21   per-adaptation evaluation statistics filled by output_transform
OUTPUT_TRANSFORM = None    # :22   set by configure_output(); the reference builds it in main() (:78-82)
DATA_DIR = None            # the reference takes it from paths.py


def configure_output(output_size=1024, data_dir=None):
    """The part of the reference's main() (:78-82) that output_transform depends on: Resize + CenterCrop + ToTensor."""
    global OUTPUT_TRANSFORM, DATA_DIR
    from torchvision import transforms
    OUTPUT_TRANSFORM = transforms.Compose([transforms.Resize(output_size), transforms.CenterCrop(output_size),
                                           transforms.ToTensor()])
    if data_dir is not None:
        DATA_DIR = data_dir


DEFAULT_TRANS = ['exposure', 'saturation', 'tone', 'color', 'contrast', 'sharp', 'blur', 'scale']


# Start value of every filter the reference can optimise (:121-209): a float for the scalar filters, a tensor for the others.
# (A factory per entry: each call of init_params must hand out fresh tensors.)
_START = {
    "gamma": lambda: 1.0, "contrast": lambda: 1.0, "saturation": lambda: 1.0,
    "sharp": lambda: 0.0, "wb": lambda: 0.0, "bright": lambda: 0.0, "exposure": lambda: 0.0, "bw": lambda: 0.0,
    "hue": lambda: 0.0,
    "blur": lambda: 1e-4,                                             # sigma ~ 0: the 25-tap Gaussian is a delta
    "tone": lambda: torch.ones(1, 8, 1),                              # identity curve, shared by the colour channels
    "color": lambda: torch.ones(3, 8, 1),                             # identity curve per colour channel
    "affine": lambda: torch.eye(2, 3),
    "scale": lambda: torch.tensor([[1.0, 1.0, 0.0, 0.0]]),            # (sx, sy, cx, cy)
}


def init_params(trans_to_apply):
    """-> (dict filter -> start value, in application order; flat start vector x0 in the same order).  Unknown names are
    skipped, as in the reference."""
    params, flat = {}, []
    for name in trans_to_apply:
        if name not in _START:
            continue
        value = _START[name]()
        params[name] = value
        flat.extend(value.reshape(-1).tolist() if torch.is_tensor(value) else [value])
    return params, torch.tensor(flat)


def initialize_parametric(image, obj_params):
    """Objective arguments + start vector for the default eight-filter chain (:212-234)."""
    params_trans, x0 = init_params(DEFAULT_TRANS)
    obj_params.update(image=image, params=params_trans)
    return x0.to(image.device), obj_params


def _param_len(template) -> int:
    return 1 if isinstance(template, float) else template.numel()


def get_params_from_vector(x, batch_size, params, input_size=480):
    """Slice the flat vector `x` back into the filter dictionary `params` (in place, in its key order) with the reference's
    shapes and clamps (:262-292): curves as [B, C, 8, 1]; affine as [B, 2, 3]; scale factors >= 1 (no black margin) and,
    for the 4-value form, the centre inside [0, input_size]; a negative contrast becomes the python float 0.0.  The clamps
    are torch ops on slices of `x`, so gradients flow exactly as in the reference."""
    start = 0
    for name in list(params.keys()):
        n = _param_len(params[name])
        piece = x[start] if n == 1 else x[start:start + n]
        start += n
        if n == 1:
            params[name] = piece
        elif name == "tone":
            params[name] = piece.view(batch_size, 1, 8, 1)
        elif name == "color":
            params[name] = piece.view(batch_size, 3, 8, 1)
        elif name == "affine":
            params[name] = piece.view(2, 3)[None].repeat(batch_size, 1, 1)
        elif name == "scale":
            if piece.size(0) == 4:
                piece = torch.cat((piece[0:2].clamp(min=1.0, max=torch.inf), piece[2:].clamp(min=0.0, max=input_size)))
            else:
                piece = piece.clamp(min=1.0, max=5.0)
            params[name] = piece.repeat(batch_size, 1)
    if params["contrast"] < 0:                      # 'contrast' is a required key, as in the reference
        params["contrast"] = 0.0
    return params


def objective_function_parametric(x_opt, image, params, clf, weight_clf, weight_dis, weight_recon, dis=None,
                                  target=None):                                          # :237-259
    params_x = get_params_from_vector(x_opt, 1, params, image.size(2))
    outputs = apply_params(image, params_x)
    loss = weight_clf * clf(outputs[-1], target=target)
    if dis is not None and weight_dis > 0:
        loss = loss - weight_dis * dis(image)
    if weight_recon > 0:                            # CLIP reconstruction term; needs the third-party image tower (see there)
        from .baselines.optimize_image import compute_clip_loss
        loss = loss + weight_recon * compute_clip_loss(image, outputs[-1])
    return loss


def output_transform(image, x_opt, obj_params, eval_params, adaptation, image_path):      # :295-312
    """Evaluate the optimised parameters on the working-size image, then re-render the edit on the full-resolution file
    (every filter is resolution-independent; `scale` centres are rescaled by get_params_from_vector's input_size)."""
    from .baselines.run_img_trans import compare_emotions
    from .baselines.utils import check_init_stats_adapt
    params_x = get_params_from_vector(x_opt, 1, obj_params["params"], image.size(2))
    outputs = apply_params(image, params_x)
    if "scale" in params_x:
        print(f"optimized scale: {params_x['scale'].tolist()}")

    check_init_stats_adapt(STATS, adaptation)
    compare_emotions(obj_params["clf"], image, outputs[-1], eval_params["emotion_type_labels"], STATS[adaptation])

    if OUTPUT_TRANSFORM is None:
        raise _lib.RgieError("output_transform: call configure_output(output_size, data_dir) first (reference main() :78-82)")
    import PIL.Image
    path = f"{DATA_DIR}/images/{image_path[0]}"
    print(path)
    input_image = PIL.Image.open(path)
    if input_image.mode != "RGB":
        input_image = input_image.convert('RGB')
    input_image_tensor = OUTPUT_TRANSFORM(input_image).unsqueeze(0).to(image.device)
    output_image_tensor = apply_params(input_image_tensor, params_x)[-1]
    return input_image_tensor, output_image_tensor
